"""The built library is Blackwell-native: its SASS holds tcgen05 MMAs (UTCHMMA), TMEM loads (LDTM), bulk-TMA copies (UBLKCP),
mbarrier ops (SYNCS), cluster barriers (UCGABAR_*) and the programmatic-dependent-launch pair (PREEXIT / ACQBULK), and no
warp-level mma.sync / wgmma fallbacks (HMMA / HGMMA).  Runs without a GPU (cuobjdump reads the cubin inside the .so);
scripts/sass_grep.sh writes the same census to profiles/r02_sass_grep.txt."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'pyprob_b200', 'lib', 'libpyprob_b200.so')


@pytest.mark.skipif(shutil.which('cuobjdump') is None, reason='cuobjdump not on PATH')
@pytest.mark.skipif(not os.path.exists(LIB), reason='library not built')
def test_library_sass_is_tcgen05_tma_and_pdl():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, timeout=600).stdout
    assert 'sm_100a' in sass or 'SM100' in sass.upper()

    def count(mnemonic):
        return len(re.findall(r'[^A-Z]' + mnemonic + r'[. ]', sass))
    for m in ('UTCHMMA', 'LDTM', 'UTCBAR', 'UBLKCP', 'SYNCS', 'UCGABAR_ARV', 'UCGABAR_WAIT', 'PREEXIT', 'ACQBULK'):
        assert count(m) > 0, m
    for m in ('HMMA', 'HGMMA'):
        assert count(m) == 0, m
    # the persistent and the cluster GEMM kernels are in the build
    assert 'k_grouped_persistent' in sass and 'k_lstm_cluster' in sass and 'k_cluster' in sass
