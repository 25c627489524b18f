"""Seeded synthetic frames and weights.

Nothing here touches the reference or the oracle: it only manufactures inputs
of the shapes the LAV frame path consumes (3 RGB cameras 288x256, a 40k-point
LiDAR sweep, ego poses for sweep stacking) and deterministic state_dict
contents for models whose released ``.th`` weights are git-LFS pointers.

All generators are CPU, seeded, and independent of module construction order,
so the authoring container, the GPU box and the golden fixtures agree.
"""
import hashlib
import math

import numpy as np
import torch

SEED = 2021  # reference default seed: lav/train_full_v2.py:68

RGB_H, RGB_W = 288, 256          # team_code_v2/lav_agent.py:53-55
SWEEP_POINTS = 40_000            # CARLA 600k pts/s, 2 ticks @20 Hz
CAMERA_YAWS = (-60, 0, 60)       # team_code_v2/lav_agent.py:35
CAMERA_X, CAMERA_Z = 1.5, 2.4    # team_code_v2/config.yaml:4-5


def _gen(seed, tag=""):
    h = hashlib.sha256(f"{seed}:{tag}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def lidar_sweep(n=SWEEP_POINTS, seed=SEED, tag="sweep", mode="carla"):
    """(n,4) fp32 x,y,z,intensity in the LiDAR frame (sensor at z=0).

    mode 'carla'   : r~U(2,85), az~U(-pi,pi), el~U(-30deg,+10deg), ground clip at z=-2.4
    mode 'uniform' : uniform in the BEV window (worst-case pillar count)
    mode 'adversarial': all points in 16 cells (atomic contention)
    """
    g = _gen(seed, tag + mode)
    if mode == "carla":
        r = torch.rand(n, generator=g) * 83.0 + 2.0
        az = (torch.rand(n, generator=g) * 2 - 1) * math.pi
        el = torch.rand(n, generator=g) * math.radians(40.0) - math.radians(30.0)
        down = el < 0
        rmax = torch.where(down, CAMERA_Z / torch.sin(-el).clamp_min(1e-3), torch.full_like(r, 1e9))
        r = torch.minimum(r, rmax)
        x = r * torch.cos(el) * torch.cos(az)
        y = r * torch.cos(el) * torch.sin(az)
        z = r * torch.sin(el)
    elif mode == "uniform":
        x = torch.rand(n, generator=g) * 79.9 - 9.95
        y = torch.rand(n, generator=g) * 79.9 - 39.95
        z = torch.rand(n, generator=g) * 3.0 - 2.4
    elif mode == "adversarial":
        cell = torch.randint(0, 16, (n,), generator=g)
        x = 10.0 + (cell % 4).float() * 0.25 + torch.rand(n, generator=g) * 0.24
        y = -2.0 + (cell // 4).float() * 0.25 + torch.rand(n, generator=g) * 0.24
        z = torch.rand(n, generator=g) * 2.0 - 2.0
    else:
        raise ValueError(mode)
    inten = torch.rand(n, generator=g)
    return torch.stack([x, y, z, inten], dim=1).float().contiguous()


def rgb_frames(seed=SEED, tag="rgb", smooth=False, n_cam=3, h=RGB_H, w=RGB_W):
    """(n_cam,h,w,3) uint8 RGB.  smooth=True gives low-pass noise so a trained
    segmenter's softmax is not degenerate."""
    g = _gen(seed, tag + str(smooth))
    if not smooth:
        return torch.randint(0, 256, (n_cam, h, w, 3), generator=g, dtype=torch.int64).to(torch.uint8)
    low = torch.rand(n_cam, 3, h // 16 + 1, w // 16 + 1, generator=g)
    img = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=True)
    img = img + 0.05 * torch.randn(n_cam, 3, h, w, generator=g)
    return (img.clamp(0, 1) * 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def sem_probs(seed=SEED, tag="sem", n_cam=3, c=5, h=RGB_H, w=RGB_W):
    """softmax of N(0,1) logits, (n_cam,c,h,w) fp32 (paint-only tests)."""
    g = _gen(seed, tag)
    return torch.softmax(torch.randn(n_cam, c, h, w, generator=g), dim=1).contiguous()


def ego_motion(n_sweeps=3, seed=SEED, tag="ego"):
    """poses (loc (n,2) fp64, ori (n,) fp64) for the stacked sweeps; index 0 = current."""
    g = _gen(seed, tag)
    loc = (torch.rand(n_sweeps, 2, generator=g, dtype=torch.float64) * 4 - 2)
    ori = (torch.rand(n_sweeps, generator=g, dtype=torch.float64) * 0.4 - 0.2)
    loc[0] = 0
    return loc.numpy(), ori.numpy()


def painted_sweep(n=SWEEP_POINTS, seed=SEED, tag="painted"):
    """(n,8) fp32: xyzI + 4 painted class probabilities (zeros for ~40% of points)."""
    g = _gen(seed, tag + "p")
    pts = lidar_sweep(n, seed, tag)
    sem = torch.rand(n, 4, generator=g) * (torch.rand(n, 1, generator=g) > 0.4)
    return torch.cat([pts, sem], dim=1).contiguous()


def stacked_lidar(n_per_sweep=SWEEP_POINTS, n_sweeps=3, seed=SEED, tag="stack"):
    """(P,11) fp32 stacked, ego-motion compensated sweeps with a one-hot time channel
    (same layout as lav_agent_fast.get_stacked_lidar, lav_agent_fast.py:363-383)."""
    loc, ori = ego_motion(n_sweeps, seed, tag)
    out = []
    for i in range(n_sweeps):
        s = painted_sweep(n_per_sweep, seed, f"{tag}{i}")
        d = ori[i] - ori[0]
        R = torch.tensor([[math.cos(d), math.sin(d), 0], [-math.sin(d), math.cos(d), 0], [0, 0, 1]], dtype=torch.float32)
        xyz = s[:, :3] @ R
        c0, s0 = math.cos(ori[0]), math.sin(ori[0])
        dl = (loc[i] - loc[0]) @ np.array([[c0, -s0], [s0, c0]])
        xyz[:, 0] += float(dl[0])
        xyz[:, 1] += float(dl[1])
        t = torch.zeros(n_per_sweep, n_sweeps)
        t[:, i] = 1
        out.append(torch.cat([xyz, s[:, 3:], t], dim=1))
    return torch.cat(out).contiguous()


# --------------------------------------------------------------------------- weights

def fill_state_dict_(sd, seed=SEED):
    """Deterministically overwrite every tensor of a state_dict in place.

    Values depend only on (seed, key, shape) so any process that builds a module
    with the same keys gets bit-identical weights.  BatchNorm statistics are
    non-trivial on purpose (identity BN hides folding bugs).
    """
    for k, v in sd.items():
        g = _gen(seed, "w:" + k)
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            v.fill_(100)
        elif leaf == "running_mean":
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif leaf == "running_var":
            v.copy_(torch.rand(v.shape, generator=g) * 1.0 + 0.5)
        elif v.dim() == 0:
            pass  # scalar buffers/params (offset_x/offset_y) keep constructor values
        elif v.dim() == 1 and leaf == "weight" and ".bn2." in k:
            # last norm of a residual branch (ResNet BasicBlock / ERFNet nb1d): keep the branch
            # small so stacked residual adds stay O(1) with un-trained weights
            v.copy_(torch.rand(v.shape, generator=g) * 0.2 + 0.2)
        elif v.dim() == 1 and leaf == "weight":      # norm scale
            v.copy_(torch.rand(v.shape, generator=g) * 1.0 + 0.5)
        elif v.dim() == 1:                            # any bias
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith(".q") or leaf == "q":
            v.copy_(torch.randn(v.shape, generator=g))
        else:
            # conv / linear / gru weights.  ConvTranspose weights are (Cin,Cout,kh,kw);
            # Kaiming-ish scale on the contracted fan so deep ReLU stacks stay O(1).
            if v.dim() >= 3:
                rf = int(np.prod(v.shape[2:]))
                fan = v.shape[1] * rf
                if ".upconv" in k or ".output_conv" in k or _is_transposed_key(k):
                    fan = v.shape[0] * max(1, rf // 4)
            else:
                fan = v.shape[-1]
            std = math.sqrt(2.0 / max(fan, 1))
            if "gru" in k:
                std = 1.0 / math.sqrt(v.shape[-1])
            v.copy_(torch.randn(v.shape, generator=g) * std)
    return sd


def _is_transposed_key(k):
    # Head.net.3 (lidar.py:155), ERFNet UpsamplerBlock.conv / decoder.output_conv (erfnet.py:102,122)
    return k.endswith("_head.net.3.weight") or ("decoder.layers.0.conv" in k) or ("decoder.layers.3.conv" in k)


def loss_block_inputs(B=4, K=5, seed=SEED, num_cmds=6, num_plan=20, num_plan_iter=5):
    """Seeded stand-ins for everything the loss block of LAV.train_lidar consumes (lav/lav_final_v2.py:177-225): the five
    LiDARModel outputs, the eleven UniPlanner outputs and the targets.  Used by oracle/pin_against_reference.py (which feeds
    them to the REFERENCE's own train_lidar through stub sub-models) and by the tests (which feed lav_b200.train)."""
    g = _gen(seed, f"lossblock{B}")
    r = lambda *s: torch.randn(*s, generator=g)
    u = lambda *s: torch.rand(*s, generator=g)
    outs = (r(B, 384, 8, 8), r(B, 2, 320, 320) * 2, u(B, 2, 320, 320) * 3, r(B, 2, 320, 320), torch.sigmoid(r(B, 3, 320, 320)))
    planner = (r(K, num_plan, 2), r(K, num_cmds, num_plan, 2), torch.sigmoid(r(K, num_cmds)), r(K, num_cmds, num_plan, 2),
               torch.sigmoid(r(K, num_cmds)), r(B, num_plan, 2), r(B, num_plan_iter, num_cmds, num_plan, 2),
               r(B, num_cmds, num_plan, 2), torch.sigmoid(r(B, num_cmds)), r(B, num_cmds, num_plan, 2),
               r(B, num_plan_iter, num_cmds, num_plan, 2))
    yy, xx = torch.meshgrid(torch.arange(320.), torch.arange(320.), indexing="ij")
    heat = torch.zeros(B, 2, 320, 320)
    for b in range(B):
        for k in range(5):
            cx, cy = (u(2) * 320).tolist()
            heat[b, k % 2] = torch.maximum(heat[b, k % 2], torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 18.0))
    targets = dict(heatmaps=heat, sizemaps=u(B, 2, 320, 320) * 3, orimaps=r(B, 2, 320, 320),
                   bev=(u(B, 9, 320, 320) > 0.7).to(torch.uint8), ego_locs=r(B, num_plan + 1, 2),
                   cmds=torch.randint(0, num_cmds, (B,), generator=g), bras=torch.tensor([0, 1, 0, 0][:B] + [0] * max(0, B - 4)))
    return outs, planner, targets
