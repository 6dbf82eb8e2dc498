"""Reading the reference's Keras weight files (`keras_model.hdf5`, explainers.py:27 `keras_model.load_weights`).

Keras stores every weight tensor as an HDF5 dataset named `<layer>/<layer>/<tensor>:0` (under `model_weights/` when the
whole model was saved). The engine's container (`model.CaptioningModel`) keys tensors `<layer>/<tensor>` with the same
Keras names, so reading a file is a walk over the datasets plus that rename. h5py is needed for
the walk and is not a dependency of the package: without it `read_keras_hdf5` raises ImportError with the way out."""
import numpy as np


def _key(dataset_name):
    """'model_weights/block1_conv1/block1_conv1/kernel:0' -> 'block1_conv1/kernel'."""
    parts = [p for p in dataset_name.split("/") if p]
    if len(parts) < 2:
        return None
    tensor = parts[-1]
    if tensor.endswith(":0"):
        tensor = tensor[:-2]
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
