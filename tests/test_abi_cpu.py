"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol that
include/b200tta.h declares, and refuses to compute off a B200 (no fallback)."""
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from longcat_video_tta_b200 import _build, _lib
    _build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = (ROOT / "include" / "b200tta.h").read_text()
    names = sorted(set(re.findall(r"\b(b200tta_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200tta.h but not exported by libb200tta.so"
    from longcat_video_tta_b200._lib import SIGNATURES
    extra = {"b200tta_last_error", "b200tta_launch_count"}
    assert set(SIGNATURES) | extra == set(names), set(names) ^ (set(SIGNATURES) | extra)


def test_struct_layouts_match_header(tmp_path):
    """sizeof() of every ABI struct as gcc sees the header == the ctypes mirror in _lib.py"""
    import ctypes as C
    import subprocess
    from longcat_video_tta_b200._lib import GemmSeg, GemmEpi, AttnSeg, TensorDesc
    src = tmp_path / "sz.c"
    src.write_text(f'#include "{ROOT}/include/b200tta.h"\n#include <stdio.h>\n'
                   'int main(void){printf("%zu %zu %zu %zu", sizeof(b200tta_gemm_seg), sizeof(b200tta_gemm_epi),'
                   ' sizeof(b200tta_attn_seg), sizeof(b200tta_tensor_desc)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(GemmSeg), C.sizeof(GemmEpi), C.sizeof(AttnSeg), C.sizeof(TensorDesc)]


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(lib):
    from longcat_video_tta_b200 import _lib, ops
    assert lib.b200tta_version() >= 100
    assert lib.b200tta_selfcheck() == _lib.EARCH
    with pytest.raises(_lib.B200TTAError):
        ops.selfcheck()
    x = torch.zeros(8, 256, dtype=torch.bfloat16)
    with pytest.raises(Exception):
        ops.ln_mod_fwd(x.clone(), x, torch.zeros(1, 256), torch.zeros(1, 256), tokens_per_frame=8)


def test_sass_is_blackwell_native():
    """tcgen05 / TMA evidence in the built library (B200_PROFILING.md: UTC*MMA, UTMALDG, LDTM/STTM)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(ROOT / "longcat_video_tta_b200" / "libb200tta.so")],
                          capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"
    # the CTA-pair GEMM: cta_group::2 MMA, 2-CTA TMA loads and the multicast commit
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA.MULTICAST"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"
