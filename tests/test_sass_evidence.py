"""CPU: the built library really contains the Blackwell-native instruction streams the design claims.

`cuobjdump -sass` of torched_impala_b200/lib/libimpala_b200.so (sm_100a) is scanned per kernel for the
SASS mnemonics of tcgen05.mma (UTCHMMA), tcgen05.ld / st (LDTM / STTM), tcgen05.commit (UTCBAR), TMA
bulk copies (UBLKCP), the L2 bulk prefetch (UBLKPF), mbarriers (SYNCS) and packed fp32 math (FFMA2):
a kernel that silently fell back to FFMA / mma.sync code would fail here before it ever reaches a GPU.
(`scripts/sass_histogram.py` writes the full table to profiles/.)
"""
import os
import re
import shutil
import subprocess

import pytest

from torched_impala_b200 import _cabi

WANT = {
    # kernel name fragment -> mnemonics that must appear in it
    "mlp_fwd_tc_pair_kernel": ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "FFMA2"),
    "mlp_bwd_tc_pair_kernel": ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "FFMA2"),
    "mlp_fwd_tcw_kernel": ("UTCHMMA", "UTCBAR", "LDTM", "UBLKPF", "SYNCS", "FFMA2"),
    "mlp_bwd_tcw_kernel": ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKPF", "SYNCS", "FFMA2"),
}


@pytest.fixture(scope="module")
def sass_by_kernel():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(_cabi.LIB_PATH):
        pytest.fail(f"{_cabi.LIB_PATH} has not been built")
    out = subprocess.run([exe, "-sass", _cabi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = set()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur:
            kernels[cur].add(m.group(1))
    return kernels


@pytest.mark.parametrize("fragment", sorted(WANT))
def test_kernel_contains_blackwell_instructions(sass_by_kernel, fragment):
    hits = {name: ops for name, ops in sass_by_kernel.items() if fragment in name}
    assert hits, f"no kernel named *{fragment}* in the library"
    for name, ops in hits.items():
        missing = [w for w in WANT[fragment] if w not in ops]
        assert not missing, (name, missing)


def test_library_is_sm100a_only(sass_by_kernel):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
