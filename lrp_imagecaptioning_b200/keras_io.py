"""Reading the reference's Keras weight files (`keras_model.hdf5`, explainers.py:27 `keras_model.load_weights`).

Keras stores every weight tensor as an HDF5 dataset named `<layer>/<layer>/<tensor>:0` (under `model_weights/` when the
whole model was saved). The engine's container (`model.CaptioningModel`) keys tensors `<layer>/<tensor>` with the same
Keras names, so reading a file is a walk over the datasets plus that rename. h5py is needed for
the walk and is not a dependency of the package: without it `read_keras_hdf5` raises ImportError with the way out."""
import numpy as np


_WRAPPERS = ("external_attention_rnn_wrapper", "external_bottom_up_attention")   # models/model.py:470, 606 (+ Keras "_N" suffix)


def _key(dataset_name):
    """HDF5 dataset path -> the container's '<layer>/<tensor>' key.

    Plain layers (also nested in a sub-model): 'model_weights/vgg16/block1_conv1/kernel:0' -> 'block1_conv1/kernel'.
    The two attention wrappers name their own tensors '<layer>_<tensor>' (models/model.py:555-570, 702-724:
    `name="{}_Wv".format(self.name)`) and hold the wrapped LSTM's kernel / recurrent_kernel / bias (built inside the
    wrapper's scope or under 'lstm_N/'), so Keras writes
        'model_weights/<layer>/<layer>/<layer>_Wv:0', '.../<layer>/<layer>/kernel:0' or '.../<layer>/lstm_1/kernel:0'
    -> '<layer>/Wv', '<layer>/kernel'."""
    parts = [p for p in dataset_name.split("/") if p]
    if len(parts) < 2:
        return None
    tensor = parts[-1]
    if tensor.endswith(":0"):
        tensor = tensor[:-2]
    for p in parts[:-1]:
        if p.startswith(_WRAPPERS):
            if tensor.startswith(p + "_"):
                tensor = tensor[len(p) + 1:]
            return p + "/" + tensor
    return parts[-2] + "/" + tensor


def read_keras_hdf5(path):
    """{'<layer>/<tensor>': float32 array} for every dataset of a Keras weight / model file."""
    try:
        import h5py
    except ImportError as e:    # pragma: no cover - depends on the environment
        raise ImportError("reading %r needs h5py; convert the file where h5py is available with "
                          "`python -m lrp_imagecaptioning_b200.keras_io in.hdf5 out.npz` and load the .npz" % path) from e
    out = {}

    def visit(name, obj):
        if hasattr(obj, "shape") and hasattr(obj, "dtype"):
            k = _key(name)
            if k is not None:
                out[k] = np.asarray(obj, dtype=np.float32)
    with h5py.File(path, "r") as f:
        f.visititems(visit)
    if not out:
        raise ValueError("%r holds no weight datasets" % path)
    return out


def convert(path_in, path_out):
    """Keras `.hdf5` -> the package's `.npz` (same tensor names)."""
    w = read_keras_hdf5(path_in)
    np.savez(path_out, **w)
    return sorted(w)


if __name__ == "__main__":
    import sys
    if len(sys.argv) != 3:
        raise SystemExit("usage: python -m lrp_imagecaptioning_b200.keras_io keras_model.hdf5 weights.npz")
    for name in convert(sys.argv[1], sys.argv[2]):
        print(name)
