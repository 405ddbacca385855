"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/pairec_gpu.h declares, refuses
to run without a GPU (no CPU fallback), and its host-only entries match the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pairec_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(prg_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pairec_b200.binding import EXPORTS, load_library
    lib = load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pairec_gpu.h but not exported"
    assert set(EXPORTS) <= set(names)


def test_product_does_not_link_or_import_the_oracle():
    # the oracle is test infrastructure: nothing under pairec_b200/ may reference it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pairec_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_init_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pairec_b200 import Engine, PrgError
    with pytest.raises(PrgError) as e:
        Engine(0)
    assert e.value.code == 6 and "no CPU fallback" in str(e.value)


def test_host_only_entries_match_oracle(oracle_lib):
    from pairec_b200.binding import lookup, sort_desc_host
    rng = np.random.default_rng(0)
    for n in (0, 1, 13, 50, 1000):
        s = np.round(rng.random(n), 2)   # ties
        assert (sort_desc_host(s) == oracle_lib.go_sort(s)).all()
    v = rng.random(10)
    pres = (rng.random(10) > 0.5).astype(np.uint8)
    assert (lookup(v, pres) == oracle_lib.lookup(v, pres)).all()


def test_error_strings_are_thread_local_and_null_safe():
    from pairec_b200.binding import load_library
    lib = load_library()
    assert lib.prg_sync(None) != 0
    assert b"null handle" in lib.prg_last_error()
    assert lib.prg_sort_desc_host(None, 5, None) != 0


def test_batcher_entry_points_reject_bad_arguments_without_a_gpu():
    from pairec_b200.binding import BatcherConfig, load_library
    lib = load_library()
    out = C.c_void_p(0)
    cfg = BatcherConfig(64, 0, 1000, 0)
    assert lib.prg_batcher_start(None, C.byref(cfg), C.byref(out)) != 0 and not out.value
    n = C.c_int32(0)
    assert lib.prg_batcher_recommend(None, None, None, None, C.byref(n)) != 0
    assert lib.prg_batcher_stats(None, None, None, None) != 0
    lib.prg_batcher_stop(None)   # no-op


def test_every_symbol_the_python_front_ends_call_is_exported():
    """binding.py / plugin.py are ctypes: a misspelt entry point would only fail when its line runs (on the GPU box)."""
    import ctypes
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for py, so, prefix in (("binding.py", "libpairec_gpu.so", "prg_"), ("plugin.py", "libpairec_host.so", "ph_")):
        src = open(os.path.join(root, "pairec_b200", py)).read()
        names = set(re.findall(r"\b(?:_lib|lib|self\._lib)\.(" + prefix + r"[a-z0-9_]+)", src))
        names |= set(re.findall(r'"(' + prefix + r'[a-z0-9_]+)"', src))
        assert len(names) >= 10, (py, names)
        lib = ctypes.CDLL(os.path.join(root, "pairec_b200", so))
        missing = [n for n in sorted(names) if not hasattr(lib, n)]
        assert not missing, f"{py} calls symbols {so} does not export: {missing}"
