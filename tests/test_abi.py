"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol include/*.h declares,
and its host-only pieces behave (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hashgan_b200 import _native
from tests import helpers


def _declared_symbols():
    text = open(os.path.join(helpers.ROOT, "include", "hashgan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.lib()
    names = _declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/hashgan_b200.h but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature in hashgan_b200/_native.py"


def test_header_declares_every_exported_symbol():
    """The other direction: nothing is exported (and nothing is bound in _native.py) that include/hashgan_b200.h does not declare."""
    import shutil
    import subprocess

    declared = set(_declared_symbols())
    assert set(_native.SIGNATURES) <= declared, sorted(set(_native.SIGNATURES) - declared)
    if shutil.which("nm") is None:
        pytest.skip("nm not available")
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("hg_")}
    assert exported, "no hg_* symbols exported"
    assert exported <= declared, f"exported but not declared in the header: {sorted(exported - declared)}"


def test_version_and_word_counts():
    lib = _native.lib()
    assert lib.hg_version() >= 100
    assert [lib.hg_code_words(b) for b in (1, 32, 33, 48, 64, 96, 128, 129, 256)] == [1, 1, 2, 2, 2, 3, 4, 8, 8]
    assert lib.hg_code_words(0) == 0 and lib.hg_code_words(257) == 0
    assert [lib.hg_label_words(L) for L in (1, 10, 32, 33, 81, 128)] == [1, 1, 1, 2, 3, 4]
    assert lib.hg_label_words(129) == 0
    # row stride: code words + label words, rounded so the code words stay 8/16-byte aligned
    assert [lib.hg_row_words(b, L) for b, L in ((32, 10), (48, 10), (64, 10), (96, 10), (128, 81), (256, 10))] == [2, 4, 4, 4, 8, 12]
    with pytest.raises(ValueError):
        _native.code_words(300)


def test_argument_errors_do_not_need_a_gpu():
    lib = _native.lib()
    assert lib.hg_hamming_map_workspace_bytes(10, 100, 64, 10, 101) == 0  # R > ndb
    assert lib.hg_hamming_map_workspace_bytes(10, 100, 300, 10, 10) == 0  # unsupported b
    assert lib.hg_hamming_map_workspace_bytes(10000, 1000000, 64, 10, 5000) > 0
    rc = lib.hg_hamming_map(None, 4, None, 10, 64, 10, 11, 0, None, None, None, None, None, 0, None)
    assert rc == _native.HG_ERANGE and b"exceeds" in lib.hg_last_error()
    rc = lib.hg_pack_rows(None, 999, None, 8, 4, 999, 10, None, None, None)
    assert rc == _native.HG_EINVAL
    with pytest.raises(_native.HgError):
        _native.check(rc)


def test_mean_matches_numpy_bitwise():
    """hg_mean_ap_host restates lib/metric.py:24 (np.mean over the kept queries) including NumPy's pairwise sum."""
    lib = _native.lib()
    rng = np.random.default_rng(5)
    for n in (0, 1, 5, 8, 9, 100, 128, 129, 1000, 4097, 10000):
        ap = rng.random(n)
        ap[rng.random(n) < 0.1] = np.nan
        out, used = C.c_double(), C.c_int64()
        assert lib.hg_mean_ap_host(ap.ctypes.data, n, C.byref(out), C.byref(used)) == 0
        kept = ap[~np.isnan(ap)]
        assert used.value == kept.size
        if kept.size == 0:
            assert np.isnan(out.value)
        else:
            assert out.value == float(np.mean(kept)), n


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/hashgan_b200.h must be usable from C (the drop-in boundary is a C ABI): compile a C99 translation unit that
    includes it with -pedantic, link it against libhashgan_b200.so and call a host-only entry point."""
    import shutil
    import subprocess

    from hashgan_b200 import _native

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "hashgan_b200.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '    if (hg_code_words(64) != 2 || hg_label_words(81) != 3 || hg_row_words(64, 10) != 4 || hg_row_words(64, 81) != 8) return 1;\n'
                   '    if (hg_crc32c("123456789", 9, 0) != 0xE3069283u) return 2;\n'
                   '    printf("%d\\n", hg_version());\n    return 0;\n}\n')
    exe = tmp_path / "abi"
    lib_dir = os.path.dirname(_native.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(helpers.ROOT, "include"), str(src), "-o", str(exe),
                    "-L", lib_dir, "-lhashgan_b200", f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip()
    assert int(out) >= 100


def test_row_layout_rules():
    """hg_row_words: rows hold the code and label words, code words stay vector-loadable, and every 33..128-bit hash length
    gets 16- or 32-byte rows (what the tensor-core select stages by TMA) for every supported label width."""
    lib = _native.lib()
    for b in list(range(1, 257)):
        W = lib.hg_code_words(b)
        assert W == ((b + 31) // 32 if b <= 128 else 8)
        for L in (1, 10, 32, 33, 64, 65, 81, 96, 128):
            LW, Wr = lib.hg_label_words(L), lib.hg_row_words(b, L)
            assert LW == (L + 31) // 32 and Wr >= W + LW
            if W == 2:
                assert Wr % 2 == 0
            if W in (4, 8):
                assert Wr % 4 == 0
            if 32 < b <= 128:
                assert Wr in (4, 8) and lib.hg_select_backend(b, L) == (64 if b <= 64 else 128)
            elif b > 128:
                assert Wr == 12 and lib.hg_select_backend(b, L) == 256        # two 128-byte K blocks
            else:  # short codes: 8- or 16-byte rows go through TMA (2-word rows as pairs), 3- and 5-word rows stay on POPC
                assert lib.hg_select_backend(b, L) == (32 if Wr in (2, 4) else 0)
    assert lib.hg_code_words(0) == 0 and lib.hg_code_words(257) == 0 and lib.hg_label_words(129) == 0 and lib.hg_row_words(64, 0) == 0
    # the kernel a whole problem gets: short codes leave the POPC kernel only for a sparse top-R (make_plan)
    assert lib.hg_select_backend_for(1000, 54000, 32, 10, 54000) == 1          # C1: MAP_R == DB_SIZE -> dense walk, no selection
    assert lib.hg_select_backend_for(1000, 54000, 32, 10, 20000) == 0          # dense-ish top-R on short codes: POPC select
    assert lib.hg_select_backend_for(1000, 1000000, 32, 10, 5000) == 32
    assert lib.hg_select_backend_for(10000, 1000000, 64, 10, 5000) == 64
    # which tensor-core select: the queued epilogue while the top-R is sparse on 33..64-bit codes in 4-word rows (C4), the
    # tile-walking one for a dense top-R (C2), 128-bit codes (C5); no select at all on the dense walk (C1)
    assert lib.hg_select_queued_for(10000, 1000000, 64, 10, 5000) == 1
    assert lib.hg_select_queued_for(1250, 1000000, 64, 10, 5000) == 1              # one rank's share at N = 8 (strong scaling)
    assert lib.hg_select_queued_for(10000, 100000, 48, 10, 5000) == 0
    assert lib.hg_select_queued_for(5000, 2000000, 128, 81, 5000) == 0
    assert lib.hg_select_queued_for(1000, 54000, 32, 10, 54000) == 0
    assert lib.hg_select_queued_for(10, 100, 64, 10, 101) == -1                    # R > ndb: no plan
    assert lib.hg_select_backend_for(100, 100000, 200, 10, 5000) == 256
    assert lib.hg_select_backend_for(100, 1000, 64, 10, 5000) == -1             # R > ndb: no plan
