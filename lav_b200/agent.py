"""Frame pipeline: the GPU side of LAVAgent.run_step (team_code_v2/lav_agent_fast.py:205-360) for B
independent agent ticks at once — seg -> paint -> sweep stack -> LiDAR model -> detection decode ->
UniPlanner -> brake.  The CARLA shell (sensors, EKF, PID, route commands) stays where it is; it hands
this class the tensors it already holds and gets back what it feeds to the PID / safety logic.

State per agent (``SweepHistory``): the FIFO of painted sweeps + ego poses that get_stacked_lidar reads
(lav_agent_fast.py:363-383), resident on the device.
"""
import math
from collections import deque

import numpy as np
import torch

from . import ops
from .model_inference import InferModel

NUM_REPEAT = 4
GAP = NUM_REPEAT + 1          # lav_agent_fast.py:31-32
NUM_FRAME_STACK = 2           # team_code_v2/config.yaml: num_frame_stack
MAX_LIDAR_POINTS = 120000


class SweepHistory:
    """FIFO of fused sweeps (N,8) with the ego pose they were taken at (lav_agent_fast.py:267-274)."""

    def __init__(self, num_frame_keep=(NUM_FRAME_STACK + 1) * GAP):
        self.lidars, self.locs, self.oris = deque(), deque(), deque()
        self.keep = num_frame_keep

    def push(self, fused, loc, ori):
        self.lidars.append(fused)
        self.locs.append(np.asarray(loc, dtype=np.float64))
        self.oris.append(float(ori))
        if len(self.lidars) > self.keep:
            self.lidars.popleft(); self.locs.popleft(); self.oris.popleft()

    def selected(self):
        """sweeps get_stacked_lidar picks: t, t-GAP, t-2*GAP ... (newest first)."""
        idx = list(range(len(self.lidars) - 1, -1, -GAP))[:NUM_FRAME_STACK + 1]
        return [(self.lidars[t], self.locs[t], self.oris[t]) for t in idx]


def stack_into(dst, sweeps, roof_filter=False):
    """get_stacked_lidar + move_lidar_points (lav_agent_fast.py:363-383,547-565) into rows of dst (P,11).
    sweeps: [(fused (n,8), loc, ori)] newest first.  Returns the number of rows written."""
    loc0, ori0 = sweeps[0][1], sweeps[0][2]
    c0, s0 = math.cos(ori0), math.sin(ori0)
    row = 0
    for i, (s, loc, ori) in enumerate(sweeps):
        d = ori - ori0
        R = np.array([[math.cos(d), math.sin(d), 0], [-math.sin(d), math.cos(d), 0], [0, 0, 1]])
        dl = (loc - loc0) @ np.array([[c0, -s0], [s0, c0]])
        n = s.shape[0]
        ops.stack_sweep(s, R, dl[0], dl[1], i, NUM_FRAME_STACK + 1, dst[row:row + n], roof_filter=roof_filter)
        row += n
    return row


class FramePipeline:
    def __init__(self, seg_model, lidar_model, uniplanner, bra_model, camera_x=1.5, camera_z=2.4, device=torch.device("cuda"),
                 precision="bf16"):
        self.device = device
        self.seg_model = seg_model.to(device).eval()
        self.bra_model = bra_model.to(device).eval() if bra_model is not None else None
        self.infer_model = InferModel(lidar_model.to(device).eval(), uniplanner.to(device).eval(), camera_x, camera_z, device)
        self.set_precision(precision)

    def set_precision(self, precision):
        self.precision = precision
        if precision == "fp32":   # exact path: keep cuDNN (PyTorch heads) out of TF32 as well
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
        self.seg_model.set_precision(precision)
        self.infer_model.lidar_model.set_precision(precision)
        dt = torch.bfloat16 if precision == "bf16" else torch.float32
        emb = self.infer_model.uniplanner.lidar_conv_emb
        emb.to(dt).to(memory_format=torch.channels_last)
        if self.bra_model is not None:
            self.bra_model.conv_backbone.to(dt).to(memory_format=torch.channels_last)
            self.bra_model.attn1.to(dt); self.bra_model.attn2.to(dt)
        return self

    @torch.no_grad()
    def step(self, rgbs_u8, tel_u8, lidars, histories, nxps, cmds, poses=None):
        """One tick of B agents.
        rgbs_u8 (B,3,288,256,3) uint8 RGB; tel_u8 (B,192,480,3) uint8 or None; lidars: list of (N_b,4) fp32 (roof-filtered
        current sweep, lav_agent_fast.py:233-247); histories: list of SweepHistory (updated in place); nxps (B,2); cmds (B,)
        poses: list of (loc, ori) of this tick (EKF state); default = zero motion.
        Returns dict(ego_plan_locs (B,20,2), ego_cast_locs, other_cast_locs [B], other_cast_cmds [B], pred_bra (B,), det, pred_bev)."""
        B = rgbs_u8.shape[0]
        im = self.infer_model
        # (1) semantic segmentation of the 3 cameras of every agent: logits NHWC (B*3,288,256,5)
        logits = self.seg_model.forward_nhwc(rgbs_u8.reshape(B * 3, *rgbs_u8.shape[2:]))
        logits = logits.view(B, 3, *logits.shape[1:]).permute(0, 1, 4, 2, 3)      # logical (B,3,5,H,W), channels-last storage
        # (2) paint the current sweep (softmax + background suppression fused into the gather) and push to the FIFO
        for b in range(B):
            fused = im.forward_paint(lidars[b], logits[b], logits=True)
            loc, ori = poses[b] if poses is not None else (np.zeros(2), 0.0)
            histories[b].push(fused, loc, ori)
        # (3) stack t, t-5, t-10 into the batch buffer
        sel = [h.selected() for h in histories]
        counts = [sum(s[0].shape[0] for s in ss) for ss in sel]
        P = max(counts)
        stacked = torch.empty((B, P, 8 + NUM_FRAME_STACK + 1), dtype=torch.float32, device=self.device)
        for b in range(B):
            stack_into(stacked[b], sel[b])
        # (4)-(6) LiDAR model, detections, motion forecast + plan
        out = im.forward_batch(stacked, counts, nxps, cmds)
        # (7) brake predictor on the stitched wide view + tele view (lav_agent_fast.py:257-262,318-321)
        if self.bra_model is not None and tel_u8 is not None:
            wide = rgbs_u8.permute(0, 2, 1, 3, 4).reshape(B, 288, 768, 3).permute(0, 3, 1, 2).float()
            tel = tel_u8.permute(0, 3, 1, 2).float()
            out["pred_bra"] = self.bra_model(wide.contiguous(memory_format=torch.channels_last),
                                             tel.contiguous(memory_format=torch.channels_last))
        return out
