"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/oh_whisper.h declares; PODs mirror the header; defaults are SURVEY.md Appendix A.2.  No compute calls."""
import ctypes as C
import os
import re

import pytest

from openhush_b200 import _abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    lib.build()
    return lib.load()


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "oh_whisper.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"OH_API[^;(]*?\b(oh_\w+)\s*\(", txt)))


def test_header_and_binding_agree(L):
    declared = _declared_symbols()
    assert len(declared) >= 45
    assert sorted(lib.SYMBOLS) == declared
    for name in declared:
        assert hasattr(L, name), name


def test_default_params_match_appendix_a2(L):
    for strategy in (0, 1):
        got = L.oh_full_default_params(strategy)
        exp = _abi.default_full_params(strategy)
        for name, _ in _abi.FullParams._fields_:
            assert getattr(got, name) == getattr(exp, name), (strategy, name)
    cp = L.oh_context_default_params()
    assert (cp.use_gpu, cp.gpu_device, cp.compute_type, cp.max_batch) == (1, 0, _abi.COMPUTE_F16, 1)


def test_no_cpu_fallback(L, micro_path):
    """use_gpu=false (the engine's `device = "cpu"`, src/engine/whisper.rs:156-157) must fail loudly, and so must a box
    without a usable GPU: NULL + message, which the engine maps to WhisperError::LoadFailed."""
    p = L.oh_context_default_params()
    p.use_gpu = 0
    assert not L.oh_init_from_file(micro_path.encode(), p)
    assert b"no CPU path" in L.oh_last_error()
    import torch
    if not torch.cuda.is_available():
        p.use_gpu = 1
        assert not L.oh_init_from_file(micro_path.encode(), p)
        assert L.oh_last_error()


def test_lang_ids(L):
    # order of /root/reference/src/engine/whisper.rs:627-731
    assert L.oh_lang_id(b"en") == 0 and L.oh_lang_id(b"de") == 2 and L.oh_lang_id(b"su") == 98 and L.oh_lang_id(b"xx") == -1


def test_product_never_links_the_oracle():
    out = os.popen(f"nm -D --defined-only {lib.SO_PATH}").read()
    assert "orc_" not in out
    for f in os.listdir(lib.CSRC):
        if f.endswith((".cu", ".cuh", ".h")):
            src = open(os.path.join(lib.CSRC, f)).read()
            assert "oracle/" not in src.replace("never includes anything from oracle/", "").replace("the\n// oracle under", "")


def test_rust_sys_crate_declares_every_symbol():
    """rust/ohwhisper-sys/src/lib.rs (source only: no Rust toolchain in the image) must declare exactly the header's symbols,
    and its PODs must list the header's fields in the header's order (the ctypes mirrors in _abi.py are the reference)."""
    txt = open(os.path.join(ROOT, "rust", "ohwhisper-sys", "src", "lib.rs")).read()
    rust_fns = sorted(set(re.findall(r"pub fn (oh_\w+)\s*\(", txt)))
    assert rust_fns == _declared_symbols()

    def rust_fields(name):
        body = re.search(r"pub struct " + name + r" \{(.*?)\n\}", txt, flags=re.S).group(1)
        return re.findall(r"pub (\w+):", body)

    assert rust_fields("oh_full_params") == [n for n, _ in _abi.FullParams._fields_]
    assert rust_fields("oh_context_params") == [n for n, _ in _abi.ContextParams._fields_]
    assert rust_fields("oh_timings") == [n for n, _ in _abi.Timings._fields_]
    assert rust_fields("oh_kernel_stat") == [n for n, _ in _abi.KernelStat._fields_]
    assert rust_fields("oh_preprocess_params") == [n for n, _ in _abi.PreprocessParams._fields_]
    hdr = open(os.path.join(ROOT, "include", "oh_whisper.h")).read()
    body = re.search(r"typedef struct oh_timings \{(.*?)\} oh_timings;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    hdr_fields = [f.strip() for decl in re.findall(r"(?:float|int|long long)\s+([^;]+);", body) for f in decl.split(",")]
    assert hdr_fields == [n for n, _ in _abi.Timings._fields_]
