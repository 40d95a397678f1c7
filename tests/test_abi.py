"""The C-ABI shared library loads without a GPU and exports every symbol include/ptb200.h declares; POD layouts
match the numpy/ctypes mirrors. No compute calls here."""
import ctypes
import os
import re


def test_every_declared_symbol_is_exported(ptb, root):
    header = open(os.path.join(root, "include", "ptb200.h")).read()
    declared = set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", header)) - {"ptb_progress_fn", "ptb_pass_fn"}
    bound = {name for name, _, _ in ptb._lib.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)
    lib = ctypes.CDLL(ptb._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert ptb._lib.lib.ptb_abi_version() == 2


def test_pod_layouts(ptb):
    L = ptb._lib
    assert L.sphere_dtype.itemsize == 20 and L.triangle_dtype.itemsize == 76
    assert L.material_dtype.itemsize == 28 and L.texture_dtype.itemsize == 28
    assert L.camera_dtype.itemsize == 48 and L.sky_dtype.itemsize == 12
    assert L.ray_dtype.itemsize == 32 and L.hit_dtype.itemsize == 16 and L.bvh_node_dtype.itemsize == 64
    assert ctypes.sizeof(L.RenderOpts) == 48 and L.RenderOpts.seed.offset == 32 and L.RenderOpts.row_begin.offset == 40
    assert ctypes.sizeof(L.Stats) == 12 * 8 + 7 * 8


def test_no_cpu_fallback(ptb):
    """Without a device the product path fails loudly; it never routes through the oracle."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        ptb.Context(0)
    except ptb.PtbError as e:
        assert e.code == 2 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("Context() must fail without a CUDA device")
    src = "".join(open(os.path.join(os.path.dirname(ptb._lib.__file__), f)).read()
                  for f in ("_lib.py", "backend.py", "scene.py", "multi.py", "meshgen.py", "__init__.py"))
    assert "oracle" not in src.replace("oracle and device", "")
