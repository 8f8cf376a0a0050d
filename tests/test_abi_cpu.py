"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/ddpm_ood_b200.h declares; the
ctypes mirror in ddpm_ood_b200/_lib.py binds exactly that set and its struct layouts agree with the compiled ones.
No compute entry point is called (there is no GPU here)."""
import ctypes as C
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "ddpm_ood_b200.h"


def _declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"DDPM_API\s+[\w\s\*]+?\b(ddpm_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared_symbols()
    assert len(names) >= 30
    for must in ("ddpm_unet_forward", "ddpm_unet_run_chain", "ddpm_add_noise", "ddpm_plms_update", "ddpm_clamp_mse",
                 "ddpm_lpips_forward", "ddpm_conv_forward", "ddpm_gn_finalize", "ddpm_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from ddpm_ood_b200 import _lib

    assert _lib.LIB_PATH.exists(), "build first: python -m ddpm_ood_b200.csrc.build"
    raw = C.CDLL(str(_lib.LIB_PATH))
    for name in _declared_symbols():
        assert hasattr(raw, name), f"{name} is declared in the header but not exported by the library"


def test_ctypes_mirror_binds_exactly_the_header():
    from ddpm_ood_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    L = _lib.lib()
    assert L.ddpm_abi_version() == _lib.ABI_VERSION
    assert L.ddpm_last_error() is not None


def test_struct_layouts_agree_with_the_compiled_library():
    from ddpm_ood_b200 import _lib

    sizes = [C.c_int(0) for _ in range(4)]
    _lib.lib().ddpm_struct_sizes(*[C.byref(s) for s in sizes])
    assert sizes[0].value == C.sizeof(_lib.ConvArgs)
    assert sizes[1].value == C.sizeof(_lib.UNetConfig)
    assert sizes[2].value == C.sizeof(_lib.PlmsStep)
    assert sizes[3].value == C.sizeof(_lib.OpProfile)
