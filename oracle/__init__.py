"""CPU oracle for the TomoEncoders hot path (TEST INFRASTRUCTURE, not product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  Nothing under ``tomoencoders_b200/`` imports it, and the product
path raises if its CUDA library is missing instead of falling back to anything here.

Contents
--------
``ref_loader``     imports the reference's own ``Patches`` / ``Grid`` / ``voxel_processing`` unmodified
                   (``/root/reference`` must exist; only used to pin the restatement and to generate
                   ``tests/golden`` fixtures -- it never travels to the GPU box).
``structures_np``  numpy restatement of the coordinate initialisers, gather (``extract``) and
                   scatter (``fill_patches_in_volume``) -- pinned against the reference's own code
                   and its notebook known answers.
``voxel_np``       ``_rescale_data`` / ``_find_min_max`` / clip restatement.
``nets_torch``     torch-CPU fp32/fp64 restatement of the Keras graphs (U-net, encoder, decoder,
                   ICIP CAE encoder).  PARITY UNPINNED: TensorFlow is not installable here and the
                   reference holds no golden vectors for network outputs, so this restatement follows
                   the Keras layer semantics spelled out in SURVEY.md appendix A.3.
``weights``        seeded weight generator in Keras ``get_weights()`` order / layout.
``pca_np``         exact PCA (float64 covariance + eigh, sklearn sign rule) and the literal
                   ``IncrementalPCA(2, batch_size=500)`` call of the reference.
``phantom``        seeded synthetic volumes for the BASELINE configs.
"""
