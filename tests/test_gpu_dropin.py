"""Drop-in proof: the REFERENCE's own callers, unmodified, running on the lav_b200 modules.

  * team_code_v2/lav_agent_fast.py — `LAVAgent.run_step` (:205-360) is executed from the staged reference source
    (baseline/_ref, byte copies made by oracle/stage_reference.py) with `carla` / `leaderboard` / `wandb` / `matplotlib` stubbed
    and the lav_b200 classes injected under the module names the agent imports (`models.lidar`, `models.uniplanner`,
    `models.bev_planner`, `models.rgb`, `model_inference`) — exactly the import swap INTEGRATION.md describes.
  * lav/lav_final_v2.py — `LAV.train_lidar` (:140-259) is executed unmodified with lav_b200's LiDARModel / UniPlanner as its
    `lidar_model` / `uniplanner`, including its own per-step `mot_inference` (eval() -> infer -> train()).
Both need a GPU (the lav_b200 modules have no CPU path) and the staged sources; they collect on CPU.
"""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

from lav_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu
STAGED = os.path.join(util.ROOT, "baseline", "_ref")


def _need_staged():
    if not os.path.isdir(os.path.join(STAGED, "team_code_v2")):
        pytest.skip("baseline/_ref not staged (run __graft_entry__.build() where /root/reference exists)")


class _Stubs:
    """sys.modules / sys.path edits that are undone afterwards"""

    def __init__(self, mods, paths):
        self.mods, self.paths, self.saved = mods, paths, {}

    def __enter__(self):
        for k, v in self.mods.items():
            self.saved[k] = sys.modules.get(k)
            sys.modules[k] = v
        for p in self.paths:
            sys.path.insert(0, p)
        return self

    def __exit__(self, *a):
        for k, v in self.saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for p in self.paths:
            sys.path.remove(p)
        for k in ("lav_agent_fast", "pid", "ekf", "planner", "waypointer", "point_painting"):
            sys.modules.pop(k, None)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def test_reference_run_step_on_lav_b200_modules(cuda):
    _need_staged()
    import yaml
    from lav_b200 import heads, lidar, model_inference, rgb
    from tests.test_heads_cpu import uniplanner
    sys.path.insert(0, os.path.join(util.ROOT, "oracle", "refshim"))
    import carla
    sys.path.pop(0)

    class AutonomousAgent:                       # leaderboard.autoagents.autonomous_agent.AutonomousAgent (base of LAVAgent)
        def __init__(self, *a, **k):
            pass
    models_pkg = _mod("models")
    models_pkg.__path__ = []
    stubs = {
        "carla": carla,
        "wandb": _mod("wandb", init=lambda *a, **k: None, log=lambda *a, **k: None, Video=lambda *a, **k: None),
        "matplotlib": _mod("matplotlib", cm=_mod("matplotlib.cm")), "matplotlib.cm": _mod("matplotlib.cm"),
        "leaderboard": _mod("leaderboard"), "leaderboard.autoagents": _mod("leaderboard.autoagents"),
        "leaderboard.autoagents.autonomous_agent": _mod("leaderboard.autoagents.autonomous_agent", AutonomousAgent=AutonomousAgent,
                                                        Track=types.SimpleNamespace(SENSORS="SENSORS")),
        "agents": _mod("agents"), "agents.navigation": _mod("agents.navigation"),
        "agents.navigation.local_planner": _mod("agents.navigation.local_planner", RoadOption=types.SimpleNamespace()),
        # the import swap: the names lav_agent_fast.py imports resolve to lav_b200
        "models": models_pkg,
        "models.lidar": _mod("models.lidar", LiDARModel=lidar.LiDARModel),
        "models.uniplanner": _mod("models.uniplanner", UniPlanner=heads.UniPlanner),
        "models.bev_planner": _mod("models.bev_planner", BEVPlanner=heads.BEVPlanner),
        "models.rgb": _mod("models.rgb", RGBSegmentationModel=rgb.RGBSegmentationModel, RGBBrakePredictionModel=heads.RGBBrakePredictionModel),
        "model_inference": model_inference,
    }
    with _Stubs(stubs, [os.path.join(STAGED, "team_code_v2")]):
        agent_mod = importlib.import_module("lav_agent_fast")          # the reference file itself
        assert os.path.samefile(agent_mod.__file__, os.path.join(STAGED, "team_code_v2", "lav_agent_fast.py"))
        LAVAgent = agent_mod.LAVAgent
        agent = object.__new__(LAVAgent)
        # what LAVAgent.setup (:62-165) establishes — minus wandb / torch.load of the LFS-pointer checkpoints: seeded weights instead
        for k, v in yaml.safe_load(open(os.path.join(STAGED, "team_code_v2", "config.yaml"))).items():
            setattr(agent, k, v)
        agent.device = cuda
        agent.waypointer = types.SimpleNamespace(tick=lambda gps: (None, None, types.SimpleNamespace(value=4)))       # -> cmd_value 3
        agent.planner = types.SimpleNamespace(run_step=lambda gps: (5.0, 20.0))
        lm, lsd = util.lidar_model(cuda)
        up, usd = uniplanner()
        sm, ssd = util.seg_model(cuda)
        bra = heads.RGBBrakePredictionModel([4, 6, 7, 10]).eval()
        bsd = synth.fill_state_dict_(bra.state_dict())
        bra.load_state_dict(bsd)
        agent.lidar_model, agent.uniplanner, agent.seg_model, agent.bra_model = lm, up.to(cuda), sm, bra.to(cuda)
        agent.infer_model = agent_mod.InferModel(agent.lidar_model, agent.uniplanner, agent.camera_x, agent.camera_z)
        assert isinstance(agent.infer_model, model_inference.InferModel)
        agent.ekf = agent_mod.EKF(1, 1.477531, 1.393600)
        agent.ekf_initialized = False
        from collections import deque
        agent.lidars, agent.locs, agent.oris = deque([]), deque([]), deque([])
        agent.vizs, agent.num_frames, agent.prev_lidar = [], 0, None
        agent.num_frame_keep = (agent.num_frame_stack + 1) * agent_mod.GAP
        agent.turn_controller = agent_mod.PIDController(K_P=agent.turn_KP, K_I=agent.turn_KI, K_D=agent.turn_KD, n=agent.turn_n)
        agent.speed_controller = agent_mod.PIDController(K_P=agent.speed_KP, K_I=agent.speed_KI, K_D=agent.speed_KD, n=agent.speed_n)
        agent.lane_change_counter = agent.stop_counter = agent.force_move = 0
        agent.lane_changed = None
        agent.visualize = lambda *a, **k: np.zeros((8, 8, 3), np.uint8)          # cv2 drawing: out of scope

        def sensors(tick):
            rgbs = synth.rgb_frames(tag=f"drop{tick}", smooth=True).numpy()                                  # (3,288,256,3) RGB
            tel = synth.rgb_frames(tag=f"dropt{tick}", smooth=True, n_cam=1, h=288, w=480)[0].numpy()         # 288 rows, bottom 96 cropped by the agent
            bgra = lambda im: np.concatenate([im[..., ::-1], np.full(im.shape[:2] + (1,), 255, np.uint8)], -1)
            d = {f"RGB_{i}": (tick, bgra(rgbs[i])) for i in range(3)}
            d["TEL_RGB"] = (tick, bgra(tel))
            d["LIDAR"] = (tick, synth.lidar_sweep(20000, tag=f"dropl{tick}").numpy())
            d["GPS"] = (tick, np.array([0.0001 * tick, 0.0002 * tick, 0.0]))
            d["IMU"] = (tick, np.array([0, 0, 0, 0, 0, 0, 0.3 + 0.01 * tick]))
            d["EGO"] = (tick, {"speed": 3.0})
            return d
        with torch.no_grad():
            c0 = LAVAgent.run_step(agent, sensors(0), 0.0)                    # first tick only stores the sweep (:238-240)
            assert (c0.steer, c0.throttle, c0.brake) == (0.0, 0.0, 0.0)
            for t in (1, 2, 3):
                ctrl = LAVAgent.run_step(agent, sensors(t), 0.05 * t)
        assert isinstance(ctrl, carla.VehicleControl)
        assert all(np.isfinite(float(v)) for v in (ctrl.steer, ctrl.throttle, ctrl.brake))
        assert -1.0 <= ctrl.steer <= 1.0 and 0.0 <= ctrl.throttle <= 1.0
        assert len(agent.lidars) == 3 and agent.lidars[-1].shape[1] == 8 and agent.lidars[-1].is_cuda
        # the tick's painted sweep equals what the oracle of the reference computes for the same sensors
        from oracle import lav_ref as O
        d = sensors(3)
        cur = torch.cat([torch.from_numpy(d["LIDAR"][1]), torch.from_numpy(sensors(2)["LIDAR"][1])])
        cur = O.preprocess(cur)
        rgb = torch.stack([torch.from_numpy(np.ascontiguousarray(d[f"RGB_{i}"][1][..., :3][..., ::-1])) for i in range(3)]).permute(0, 3, 1, 2).float()
        with torch.no_grad():
            want = O.forward_paint(cur, torch.softmax(O.erfnet(ssd, rgb), 1), O.default_converters())
        got = agent.lidars[-1].cpu()
        assert torch.equal(got[:, :4], want[:, :4])
        assert int(((got - want).abs() > 2e-3).any(1).sum()) < 20


def test_reference_train_lidar_on_lav_b200_modules(cuda):
    _need_staged()
    import yaml
    from lav_b200.train import synthetic_train_batch
    from tests.test_heads_cpu import uniplanner
    with _Stubs({}, [os.path.join(util.ROOT, "oracle", "refshim"), STAGED]):
        for k in [k for k in sys.modules if k == "lav" or k.startswith("lav.")]:
            sys.modules.pop(k)
        LAV = importlib.import_module("lav.lav_final_v2").LAV
        RefDetLoss = importlib.import_module("lav.models.loss").DetLoss
        cfg = yaml.safe_load(open(os.path.join(STAGED, "config_v2.yaml")))
        obj = object.__new__(LAV)                         # LAV.__init__ (:19-129) minus the torch.load of LFS-pointer checkpoints
        for k, v in cfg.items():
            setattr(obj, k, v)
        obj.device, obj.distill, obj.multi_gpu = cuda, True, False
        obj.perceive_only = obj.motion_only = False
        lm, _ = util.lidar_model(cuda)
        up, _ = uniplanner()
        obj.lidar_model, obj.uniplanner = lm.train(), up.to(cuda).train()
        obj.uniplanner.bev_planner.eval()
        u = obj.uniplanner
        params = (list(u.plan_gru.parameters()) + list(u.plan_mlp.parameters()) + list(u.cast_grus_ego.parameters()) + list(u.cast_mlps_ego.parameters()) +
                  list(u.cast_grus_other.parameters()) + list(u.cast_mlps_other.parameters()) + list(u.cast_cmd_pred.parameters()) +
                  list(u.lidar_conv_emb.parameters()) + list(obj.lidar_model.parameters()))                    # lav_final_v2.py:74-86
        obj.lidar_optim = torch.optim.Adam(params, lr=1e-4)
        obj.det_criterion = RefDetLoss()
        obj.bev_center = [160.0, 280.0]
        obj.seg_mask = LAV.build_seg_mask(obj, h=320, w=320, cx=160, cy=280).to(cuda)
        obj.branch_weights = torch.tensor(cfg["branch_weights"]).float().to(cuda)
        batch = synthetic_train_batch(2, torch.device("cpu"), n_points=(20000, 30000))
        lidars, npts, heat, size, ori, bev, ego_locs, cmds, nxps, bras, locs, oris, typs = batch
        w0 = u.plan_mlp.weight.detach().clone()
        c0 = obj.lidar_model.backbone.conv1[0].weight.detach().clone()
        teacher0 = [p.detach().clone() for p in u.bev_planner.parameters()]
        names = ["hm_loss", "box_loss", "ori_loss", "seg_loss", "plan_loss", "ego_cast_loss", "other_cast_loss", "cmd_loss"]
        res = []
        for _ in range(2):                                # the reference's own method, unmodified
            res.append(LAV.train_lidar(obj, lidars, npts, heat, size, ori, bev, ego_locs, cmds, nxps, bras, locs, oris, typs, torch.tensor([6, 6])))
        assert all(np.isfinite(r[k]) for r in res for k in names)
        assert not torch.equal(w0, u.plan_mlp.weight) and not torch.equal(c0, obj.lidar_model.backbone.conv1[0].weight)
        assert all(torch.equal(a, b) for a, b in zip(teacher0, u.bev_planner.parameters()))
        assert obj.lidar_model.training and obj.uniplanner.training          # mot_inference switched to eval and back (:291-322)
        assert res[0]["ego_plan_locs"].shape == (20, 2)
