"""Multi-GPU parity inside `pytest -m gpu`: one process per GPU over NCCL (torch.distributed.run spawned from here), the
halo / north-fold exchange of tra_adv_fct against the mono-domain oracle, bit for bit -- closed, E-W cyclic, T-pivot and
F-pivot folds x 2nd / 4th order x the reference-structured, three-kernel and one-kernel schedules.  On 8 GPUs the partition
is 4x2: fold partners sit on different ranks (mpp_nfd_generic.h90:78-221, mppini.F90:1180-1240).  Skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


_port = [29533]


def _run(nproc, env_extra, timeout):
    env = dict(os.environ, **env_extra)
    _port[0] += 1                                                      # a fresh rendezvous port per launch
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(_port[0]), os.path.join(ROOT, "tests", "mgpu_check.py")]
    return subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_fct_over_nccl_all_gpus_of_the_box():
    n = torch.cuda.device_count()
    n = 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 1
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    r = _run(n, {"MGPU_ONLY_FCT": "1"}, 1500)
    sys.stdout.write(r.stdout[-6000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_CHECK PASS" in r.stdout and "MISMATCH" not in r.stdout


def test_widened_rows_over_nccl():
    n = torch.cuda.device_count()
    n = 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 1
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    r = _run(n, {}, 2400)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_CHECK PASS" in r.stdout and "MISMATCH" not in r.stdout


def test_collective_diagnostics_over_nccl():
    """glob_sum (MPI_SUMDD, lib_mpp.F90:1158-1186) and stp_ctl with ln_ctl (mpp_max / mpp_maxloc, stpctl.F90:126-129, 157-160)
    through the NCCL communicator: the same numbers on every rank as the mono-domain oracle"""
    n = torch.cuda.device_count()
    n = 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 1
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    r = _run(n, {"MGPU_ONLY_DIAG": "1"}, 600)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_CHECK PASS" in r.stdout and "MISMATCH" not in r.stdout and "glob_sum over NCCL" in r.stdout
