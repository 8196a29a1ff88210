"""``build_Unet_3D``: drop-in for /root/reference/tomo_encoders/neural_nets/Unet3D.py:199-291.
Returns a ``Model`` (keras_processor.Model) whose graph is planned and executed by libtomoenc.so."""
from .keras_processor import Model


def build_Unet_3D(n_filters=[16, 32, 64], n_blocks=3, activation="lrelu", batch_norm=True, kern_size=3,
                  kern_size_upconv=2, isconcat=None, pool_size=2, mode=None, seed=None):
    if isinstance(pool_size, int):
        pool_size = [pool_size] * n_blocks
    elif len(pool_size) != n_blocks:
        raise ValueError("list length must be equal to number of blocks")
    if isconcat is None:
        isconcat = [False] * n_blocks
    params = dict(n_filters=list(n_filters), n_blocks=n_blocks, activation=activation, batch_norm=batch_norm,
                  kern_size=kern_size, kern_size_upconv=kern_size_upconv, isconcat=list(isconcat),
                  pool_size=list(pool_size))
    return Model("unet", params, model_size=None, mode=mode, name="segmenter", seed=seed)
