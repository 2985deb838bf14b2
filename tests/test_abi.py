"""CPU checks of the drop-in boundary: libb2nn.so loads without a GPU, exports every symbol include/b2nn.h declares
(and the reference's op tokens), and fails LOUDLY — no CPU fallback — when no sm_100 device is present."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "b2nn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2nn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(b2nn):
    L = b2nn.lib()
    declared = _header_symbols()
    assert len(declared) >= 20
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/b2nn.h but not exported by libb2nn.so"
    assert sorted(b2nn.ABI_SYMBOLS) == declared


def test_reference_op_tokens_are_exported(b2nn):
    """activations.h / costfunctions.h / randinitweights.h / kernels.h names (C++ linkage, like the reference)."""
    out = subprocess.run(["nm", "-DC", b2nn.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for name in ("sigmoidGPU(float const*, float*, int)", "tanhGPU(float const*, float*, int)",
                 "sechSqGPU(float const*, float*, int)", "sigmoidGradGPU(float const*, float*, int)",
                 "softmaxGPU(float const*, float*, int, int)", "sigmoidOutputGPU(float const*, float*, int, int)",
                 "negLnMaxCostGPU(float*, float*, float*, int)", "crossEntropyCostGPU(float*, float*, float*, int)",
                 "randInitWeights1GPU(float*, int, int, int)", "randInitWeights2GPU(float*, int, int, int)",
                 "probToNumGPU(float*, float*, int, int)", "countErrorGPU(float*, float*, float*, int)"):
        assert name in out, name


def test_builtin_recognition_by_address(b2nn):
    L = b2nn.lib()
    sym = {"_Z7tanhGPUPKfPfi": (0, b2nn.ACT_TANH), "_Z10sigmoidGPUPKfPfi": (0, b2nn.ACT_SIGMOID),
           "_Z9sechSqGPUPKfPfi": (1, b2nn.ACT_TANH), "_Z14sigmoidGradGPUPKfPfi": (1, b2nn.ACT_SIGMOID),
           "_Z10softmaxGPUPKfPfii": (2, b2nn.OUT_SOFTMAX), "_Z16sigmoidOutputGPUPKfPfii": (2, b2nn.OUT_SIGMOID),
           "_Z19crossEntropyCostGPUPfS_S_i": (3, b2nn.COST_CROSSENTROPY), "_Z15negLnMaxCostGPUPfS_S_i": (3, b2nn.COST_NEGLNMAX),
           "_Z19randInitWeights1GPUPfiii": (4, b2nn.INIT_1), "_Z19randInitWeights2GPUPfiii": (4, b2nn.INIT_2)}
    for mangled, (family, kind) in sym.items():
        addr = ctypes.cast(getattr(L, mangled), ctypes.c_void_p)
        assert L.b2nn_builtin_kind(addr, family) == kind, mangled
        assert L.b2nn_builtin_kind(addr, (family + 2) % 5) == -1
    assert L.b2nn_builtin_kind(ctypes.c_void_p(0x1234), 0) == -1


def test_no_cpu_fallback_without_gpu(b2nn):
    """Without a usable sm_100 device the product path must refuse to run, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert not b2nn.device_ok()
    with pytest.raises(b2nn.B2nnError) as e:
        b2nn.Net()
    assert "rc=2" in str(e.value)     # B2NN_ERR_NO_DEVICE


def test_product_code_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under cublas-ml_b200/ or include/ may import, link or name it."""
    bad = []
    for base in ("cublas-ml_b200", "include"):
        for dirpath, _dirs, files in os.walk(os.path.join(ROOT, base)):
            if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
                continue
            for f in files:
                if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    if re.search(r"liboracle|oracle/|import oracle|from oracle", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
