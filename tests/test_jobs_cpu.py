"""Host logic of the fixed-shape pipeline (no GPU): the per-tick (agent, sweep) job table written by
StaticFramePipeline._fill_jobs — ring-buffer slot selection, relative rotation and translation — reproduces
get_stacked_lidar / move_lidar_points of the reference (team_code_v2/lav_agent_fast.py:363-383,547-565, oracle/lav_ref.py:stack_lidar)
when the records are applied the way stack_jobs_kernel applies them (csrc/paint.cu)."""
import ctypes
import types

import numpy as np
import torch

from lav_b200 import ops
from lav_b200.agent import GAP, StaticFramePipeline
from oracle import lav_ref as O


def _apply_jobs(stub):
    """numpy twin of stack_jobs_kernel: xyz' = R^T-style row products + (dx, dy), copy the other columns, one-hot time."""
    B, T, N = stub.B, stub.T, stub.N
    jobs = stub.jobs_dev.numpy().view(ops.STACK_JOB_DTYPE).reshape(B, T)
    out = np.full((B, T * N, 8 + T), np.nan, np.float32)
    for b in range(B):
        for t in range(T):
            j = jobs[b, t]
            n = int(j["n"])
            if n == 0:
                continue
            src = np.ctypeslib.as_array(ctypes.cast(int(j["src"]), ctypes.POINTER(ctypes.c_float)), shape=(n, 8))
            assert int(j["dst"]) == stub.stacked.data_ptr() + (b * T * N + t * N) * (8 + T) * 4
            R = j["R"].astype(np.float32)
            x, y, z = src[:, 0], src[:, 1], src[:, 2]
            d = out[b, t * N:t * N + n]
            d[:, 0] = z * R[6] + (y * R[3] + x * R[0]) + j["dx"]
            d[:, 1] = z * R[7] + (y * R[4] + x * R[1]) + j["dy"]
            d[:, 2] = z * R[8] + (y * R[5] + x * R[2])
            d[:, 3:8] = src[:, 3:8]
            d[:, 8:] = 0
            d[:, 8 + int(j["time_idx"])] = 1
    return out


def test_fill_jobs_matches_reference_stacking():
    B, T, N = 2, 3, 50
    KEEP = StaticFramePipeline.KEEP
    g = torch.Generator().manual_seed(7)
    stub = types.SimpleNamespace(
        B=B, T=T, N=N, KEEP=KEEP, tick=0,
        ring=torch.full((B, KEEP, N, 8), float("nan")), cur=torch.full((B, N, 8), float("nan")),
        stacked=torch.full((B, T * N, 8 + T), float("nan")),
        ring_pose=np.zeros((B, KEEP, 3)), ring_valid=np.zeros((B, KEEP), dtype=bool),
        jobs_host=torch.zeros(B * T * ops.STACK_JOB_DTYPE.itemsize, dtype=torch.uint8),
        jobs_dev=torch.zeros(B * T * ops.STACK_JOB_DTYPE.itemsize, dtype=torch.uint8))
    history = [[] for _ in range(B)]                      # per agent: list of (sweep, loc, ori), one per tick
    for tick in range(2 * GAP + 3):                       # long enough for t-5 and t-10 to exist
        poses = []
        for b in range(B):
            sweep = torch.randn(N, 8, generator=g) * torch.tensor([20., 20., 1., 1., 1., 1., 1., 1.])
            loc = np.array([0.7 * tick + b, -1.3 * tick + 0.1 * b])
            ori = 0.05 * tick * (1 if b == 0 else -1) + 0.3 * b
            stub.cur[b] = sweep
            history[b].append((sweep, loc, ori))
            poses.append((loc, ori))
        StaticFramePipeline._fill_jobs(stub, poses)
        got = _apply_jobs(stub)
        for b in range(B):
            sel = [history[b][t] for t in range(tick, -1, -GAP)][:T]                  # t, t-5, t-10 (newest first)
            ref = O.stack_lidar([s for s, _, _ in sel], [l for _, l, _ in sel], [o for _, _, o in sel]).numpy()
            rows = np.concatenate([got[b, k * N:(k + 1) * N] for k in range(len(sel))])
            assert np.isnan(got[b, len(sel) * N:]).all()                                  # sweeps that do not exist yet: no job
            np.testing.assert_allclose(rows, ref, rtol=0, atol=2e-4)
        stub.ring[:, stub.tick % KEEP] = stub.cur                                         # FIFO push, as _begin does
        stub.tick += 1
