"""Synchronisation stress test: the jittered build of the library (random sleeps in front of the
mbarrier operations of every warp role, csrc/tc_common.cuh, -DDCGRU_JITTER) runs small encoder / decoder cases 200 times
and every output and gradient must stay bit-identical to the first iteration; a lost arrival or an overtaken phase
traps instead.  One process per visible GPU (up to 2) run concurrently, which is how the round-1 deadlock showed up."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_jittered_build_200_iterations():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    lib = os.path.join(ROOT, "eeg-gnn-ssl_b200", "lib", "libdcgru_b200_jitter.so")
    assert os.path.exists(lib), "build it with __graft_entry__.build() (or eeg-gnn-ssl_b200/build.py --jitter)"
    env = dict(os.environ, DCGRU_B200_LIB="libdcgru_b200_jitter.so")
    ndev = min(torch.cuda.device_count(), 2)
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "scripts", "stress.py"), "200", str(d)], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for d in range(ndev)]
    for p in procs:
        out, _ = p.communicate(timeout=900)
        assert p.returncode == 0, out[-2000:]
        assert "mismatching tensors: 0" in out, out[-2000:]
