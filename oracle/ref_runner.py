"""Run the UNMODIFIED reference modules (staged in baseline/_ref by oracle/stage_reference.py) on one synthetic frame, on any
device — the reference arms of bench.py: `--impl reference` / `cpu_baseline` (host cores) and `gpu_reference` (the reference's own
PyTorch-CUDA path on the same B200, the denominator of the north_star's ">= 10x").

Test / measurement infrastructure: only bench.py's reference arms and the tests import this.  The frame follows
LAVAgent.run_step of team_code_v2/lav_agent_fast.py:205-360 for everything that touches a model:

    seg_model(all_rgbs) -> softmax                      (:263-264)     reference RGBSegmentationModel, eager (the agent loads a
                                                                       TorchScript trace of the same module)
    infer_model.forward_paint(cur_lidar, pred_sem)      (:266)         reference InferModel (jit-scripted converters)
    FIFO + get_stacked_lidar + move_lidar_points        (:268-277, :363-383, :547-565)   restated below in torch (that file imports
                                                                       carla/leaderboard and cannot be imported)
    infer_model(lidar_points, nxps, cmd)                (:317)         reference InferModel.forward pieces: pillar net (eager),
                                                                       jit-scripted backbone / heads / conv-embedder, det_inference,
                                                                       uniplanner_infer — fed the bench's fixed K = 3 vehicle list
    bra_model(rgbs, tel_rgbs)                           (:323)         reference RGBBrakePredictionModel, eager
Weights: lav_b200.synth.fill_state_dict_ (the released .th files are git-LFS pointers) — identical values to the lav_b200 arm.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")


def available():
    return os.path.isdir(os.path.join(STAGED, "team_code_v2", "models"))


def _import_reference():
    for p in (os.path.join(ROOT, "oracle", "refshim"), os.path.join(STAGED, "team_code_v2"), STAGED):
        if p not in sys.path:
            sys.path.insert(0, p)
    from models.lidar import LiDARModel
    from models.uniplanner import UniPlanner
    from models.bev_planner import BEVPlanner
    from models.rgb import RGBSegmentationModel, RGBBrakePredictionModel
    import model_inference as MI
    return LiDARModel, UniPlanner, BEVPlanner, RGBSegmentationModel, RGBBrakePredictionModel, MI


class ReferenceFrame:
    """reference models with seeded weights on `device`; __call__ runs one agent frame (batch 1, like the agent)."""

    def __init__(self, device, fixed_dets, jit=True):
        from lav_b200 import synth
        LiDARModel, UniPlanner, BEVPlanner, RGBSeg, RGBBra, MI = _import_reference()
        self.device = torch.device(device)
        kw = dict(pixels_per_meter=4, crop_size=96, feature_x_jitter=1.5, feature_angle_jitter=20, x_offset=0,
                  y_offset=1 + (-10) / ((70 + 10) / 2), num_cmds=6, num_plan=20, num_plan_iter=5)
        seg = RGBSeg([4, 6, 7, 10]).eval()
        lid = LiDARModel(num_input=16, num_features=[64, 64], backbone="cnn", min_x=-10, max_x=70, min_y=-40, max_y=40,
                         pixels_per_meter=4).eval()
        uni = UniPlanner(BEVPlanner(num_frame_stack=2, **kw), num_input_feature=384, **kw).eval()
        bra = RGBBra([4, 6, 7, 10], pretrained=False).eval()
        for m in (seg, lid, uni, bra):
            m.load_state_dict(synth.fill_state_dict_(m.state_dict()))
            m.to(self.device)
        self.seg_model, self.bra_model = seg, bra
        if jit:
            self.infer_model = MI.InferModel(lid, uni, 1.5, 2.4, device=self.device).to(self.device)     # jit-scripts backbone/heads/emb
        else:
            raise NotImplementedError("the agent always builds the jit-scripted InferModel (model_inference.py:20-32)")
        self.fixed_dets = list(fixed_dets)

    @staticmethod
    def _move(xyz, dloc, ori0, ori1):
        """move_lidar_points, lav_agent_fast.py:547-565"""
        dloc = np.asarray(dloc, dtype=np.float64) @ np.array([[np.cos(ori0), -np.sin(ori0)], [np.sin(ori0), np.cos(ori0)]])
        ori = ori1 - ori0
        R = torch.tensor([[np.cos(ori), np.sin(ori), 0], [-np.sin(ori), np.cos(ori), 0], [0, 0, 1]], dtype=torch.float, device=xyz.device)
        out = xyz @ R
        out[:, 0] += float(dloc[0])
        out[:, 1] += float(dloc[1])
        return out

    def _stack(self, sweeps, locs, oris):
        """get_stacked_lidar, lav_agent_fast.py:363-383 (sweeps newest first)"""
        rel = []
        for i, (s, loc, ori) in enumerate(zip(sweeps, locs, oris)):
            xyz = self._move(s[:, :3], np.asarray(loc) - np.asarray(locs[0]), oris[0], ori)
            t = torch.zeros((len(xyz), len(sweeps)), dtype=xyz.dtype, device=xyz.device)
            t[:, i] = 1
            rel.append(torch.cat([xyz, s[:, 3:], t], dim=-1))
        return torch.cat(rel)

    @torch.no_grad()
    def __call__(self, rgbs_u8, tel_u8, lidar, prev, loc, ori, nxp, cmd):
        """rgbs_u8 (3,288,256,3) u8, tel_u8 (192,480,3) u8, lidar (N,4), prev = [fused (n,8)] * 2, loc/ori poses of the 3 sweeps
        (HOST tensors, as the CARLA sensors deliver them: the H2D copies are part of the agent's frame)."""
        dev = self.device
        im = self.infer_model
        cur_lidar = torch.as_tensor(lidar, dtype=torch.float, device=dev)
        all_rgbs = rgbs_u8.permute(0, 3, 1, 2).float().to(dev)
        pred_sem = torch.softmax(self.seg_model(all_rgbs), dim=1)
        fused = im.forward_paint(cur_lidar, pred_sem)
        lidar_points = self._stack([fused] + [p.to(dev) for p in prev], loc, ori)
        nxps = torch.as_tensor(nxp, dtype=torch.float).to(dev)
        features = im.lidar_model_point_pillar([lidar_points], [len(lidar_points)])
        features = im.lidar_mode_backbone(features)
        pred_heatmaps, pred_sizemaps = im.lidar_center_head(features), im.lidar_box_head(features)
        pred_orimaps, pred_bev = im.lidar_ori_head(features), im.lidar_seg_head(features)
        det = im.det_inference(torch.sigmoid(pred_heatmaps[0]), pred_sizemaps[0], pred_orimaps[0])
        ego_embd, ego_plan_locs, ego_cast_locs, other_cast_locs, other_cast_cmds = im.uniplanner_infer(features[0], self.fixed_dets, cmd, nxps)
        rgbs = rgbs_u8.permute(1, 0, 2, 3).reshape(288, 768, 3)[None].permute(0, 3, 1, 2).float().to(dev)
        tel_rgbs = tel_u8[None].permute(0, 3, 1, 2).float().to(dev)
        pred_bra = self.bra_model(rgbs, tel_rgbs)
        return ego_plan_locs.float().cpu(), float(pred_bra), det
