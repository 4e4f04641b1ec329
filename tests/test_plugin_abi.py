"""The C ABI library loads, exports every symbol include/ngm_b200.h declares, and its NGM plugin surface
(Cookie / SetLog / SetConfig / IsAvailable / CreateAlignment / DeleteAlignment / ExternalDeleteString,
reference SWOcl_export.cpp:20-83) can be driven by a host compiled against the reference's own headers."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "nextgenmap_b200" / "libngm_b200.so"
REF_INC = Path("/root/reference/include")


def _ensure_lib():
    if not LIB.exists():
        from nextgenmap_b200 import build
        build.build()
    return LIB


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(str(_ensure_lib()))
    header = (ROOT / "include" / "ngm_b200.h").read_text()
    names = sorted(set(re.findall(r"\b(ngm_b200_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 18
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    for n in ("Cookie", "SetLog", "SetConfig", "IsAvailable", "CreateAlignment", "DeleteAlignment", "ExternalDeleteString"):
        assert hasattr(lib, n), n
    assert lib.ngm_b200_abi_version() == 1
    lib.Cookie.restype = ctypes.c_int
    assert lib.Cookie() == 0x10201130                      # IAlignment.h:30


def test_product_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nextgenmap_b200.host import CudaSW, NgmB200Error
    with pytest.raises(NgmB200Error, match="no CPU fallback"):
        CudaSW(152, 27)


def _build_client(tmp_path, use_reference_headers: bool) -> Path:
    exe = tmp_path / ("client_ref" if use_reference_headers else "client_own")
    cmd = ["g++", "-std=gnu++11", "-O1", "-w", str(ROOT / "tests" / "plugin_client.cpp"), "-ldl", "-o", str(exe)]
    cmd += ["-DUSE_REFERENCE_HEADERS", f"-I{REF_INC}"] if use_reference_headers else [f"-I{ROOT / 'include'}"]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


@pytest.mark.parametrize("use_ref", [False, True])
def test_plugin_surface_probe(tmp_path, use_ref):
    if use_ref and not REF_INC.exists():
        pytest.skip("/root/reference not present")
    exe = _build_client(tmp_path, use_ref)
    p = subprocess.run([str(exe), str(_ensure_lib()), "probe"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "cookie ok" in p.stdout


@pytest.mark.gpu
def test_plugin_vtable_calls_on_gpu(tmp_path):
    """Appendix A rows 0, 2, 6 through CreateAlignment + the IAlignment vtable."""
    exe = _build_client(tmp_path, REF_INC.exists())
    p = subprocess.run([str(exe), str(_ensure_lib()), "run"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr + p.stdout
    out = p.stdout
    assert "mode 0 pair 0 score 300 off 5 qs 0 qe 0 nm 0 id 1 cigar 30M md 30 as 30" in out
    assert "mode 0 pair 1 score 280 off 5 qs 0 qe 0 nm 1 id 0.967742 cigar 11M1D19M md 11^T19 as 30" in out
    assert "mode 0 pair 2 score 50 off 12 qs 5 qe 20 nm 0 id 1 cigar 5S5M20S md 5 as 10" in out
    assert "mode 1 pair 2 score -90 off 9 qs 0 qe 0 nm 14 id 0.533333 cigar 2M1I1M1I5M1I1M1I2M4I11M md 11G1G2A1GTA1 as 30" in out
