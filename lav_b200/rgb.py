"""RGBSegmentationModel — drop-in mirror of lav/models/rgb.py:35-45 (keys ``erfnet.{encoder,decoder}.*``).

``forward(rgb)`` takes what the reference takes (float NCHW in 0..255) and returns logits with the
reference's logical shape (N, 5, H, W), stored channels-last.  ``forward_u8`` ingests the camera
frames as the agent holds them (uint8 HWC, lav_agent.py:243) without a float round trip.
The brake model (RGBBrakePredictionModel) stays a PyTorch head: see lav_b200/heads.py.
"""
import torch
from torch import nn

from . import ops
from .capi import LavbError
from .erfnet import ERFNet, _dt


class RGBSegmentationModel(nn.Module):
    def __init__(self, seg_channels):
        super().__init__()
        self.erfnet = ERFNet(len(seg_channels) + 1)
        self.normalize = lambda x: (x / 255. - .5) * 2      # kept for API parity; fused into the ingest kernel

    def set_precision(self, precision):
        self.erfnet.precision = precision
        self.erfnet.invalidate_plan()
        return self

    def _ingest(self, rgb):
        """uint8 NHWC frames go to the ERFNet as they are (fused normalize + initial block, csrc/erf16.cu); float NCHW input
        (the reference's call form) is normalised by the ingest kernel."""
        from . import erfnet as E
        if rgb.dtype == torch.uint8 and E.FUSE_STEM and rgb.dim() == 4 and rgb.shape[3] == 3 and rgb.shape[1] % 2 == 0 and rgb.shape[2] % 2 == 0:
            return rgb.contiguous()
        return ops.rgb_normalize(rgb, _dt(self.erfnet.precision))

    def forward_nhwc(self, rgb):
        """rgb: uint8 (N,H,W,3) or float (N,3,H,W), 0..255 -> logits NHWC fp32 (N,H,W,C)."""
        if not rgb.is_cuda:
            raise LavbError("lav_b200.RGBSegmentationModel needs CUDA tensors (no CPU fallback)")
        return self.erfnet.forward_nhwc(self._ingest(rgb))

    def forward_features_nhwc(self, rgb):
        """-> (decoder features NHWC (N,H/2,W/2,16), output_conv table, n_classes): see ERFNet.forward_features_nhwc."""
        if not rgb.is_cuda:
            raise LavbError("lav_b200.RGBSegmentationModel needs CUDA tensors (no CPU fallback)")
        feat, table = self.erfnet.forward_features_nhwc(self._ingest(rgb))
        return feat, table, self.erfnet.decoder.output_conv.out_channels

    def forward_u8(self, rgb_u8_nhwc):
        return self.forward_nhwc(rgb_u8_nhwc).permute(0, 3, 1, 2)

    def forward(self, rgb):
        return self.forward_nhwc(rgb).permute(0, 3, 1, 2)
