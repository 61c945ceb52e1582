"""CPU: the C-ABI shared library loads, exports every symbol include/svo_b200.h declares,
its host-only entry points compute, and the compute path fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "svo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svo_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ooc_svo_builder_b200 import LIB_PATH, ABI_SYMBOLS
    assert os.path.exists(LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    lib = ctypes.CDLL(LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libsvo_b200.so does not export %s" % n
    assert sorted(ABI_SYMBOLS) == names, "api.ABI_SYMBOLS out of sync with include/svo_b200.h"


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "svo_b200.h")).read()
    assert "torch" not in src and "at::" not in src and "std::" not in src


def test_host_only_entry_points(oracle):
    from ooc_svo_builder_b200 import estimate_partitions, load_library
    for g, lim in [(256, 2048), (2048, 2048), (4096, 2048), (8192, 2048), (256, 3), (512, 2), (1024, 100)]:
        assert estimate_partitions(g, lim) == oracle.estimate_partitions(g, lim)
    lib = load_library()
    olib = oracle.lib()
    for v in (2.0, 1.9, 1.23456789, 0.1, 3.4028235e38, 1e-30, 7.0 / 3.0):
        assert lib.svo_text_roundtrip_float(v) == olib.svo_oracle_text_roundtrip(v)
    assert b"b200" in lib.svo_version().lower()


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ooc_svo_builder_b200 import SvoBuilder, SvoError
    with pytest.raises(SvoError) as e:
        SvoBuilder(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under the package, the tools or the public header may reference it
    for top in ("ooc_svo_builder_b200", "tools", "include"):
      for d, _, files in os.walk(os.path.join(ROOT, top)):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert ("liboracle" not in txt and "svo_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt
                        and '"oracle"' not in txt), os.path.join(d, f)
