"""CPU: the C-ABI library loads without a GPU and exports every symbol include/lrpcap.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lrpcap.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lrpcap_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    path = g.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n


def test_python_prototypes_cover_the_header():
    from lrp_imagecaptioning_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared()


def test_error_plumbing_without_gpu():
    from lrp_imagecaptioning_b200 import _lib
    lib = _lib.load()
    assert lib.lrpcap_version() >= 100
    assert lib.lrpcap_encoder_forward(None, None, 0, 0, 0.0, 0.0, 0.0, 1, None) == -1
    assert b"null handle" in lib.lrpcap_last_error()
    with pytest.raises(_lib.LrpcapError):
        _lib.check(lib.lrpcap_decoder_forward(None, None, 1, 1, None, 1, 0, -1, None))


def test_product_path_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the package may reference it."""
    pkg = os.path.join(ROOT, "lrp_imagecaptioning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
