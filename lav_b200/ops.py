"""Torch-tensor front ends of the C-ABI kernels (device pointers + current stream in, tensors out).

PyTorch is plumbing here: allocation, streams, views.  Every function launches hand-written
sm_100a kernels through lav_b200.capi; nothing falls back to torch math.
"""
import ctypes as C

import numpy as np
import torch

from . import capi
from .capi import BF16, F16, F32, ConvDesc, check, lib

_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


def h16():
    """torch dtype of the library's 16-bit storage type: float16 (fp32 accumulation, saturating stores) unless the library was
    built with -DLAVB_H16_BF16."""
    return torch.float16 if capi.h16_code() == F16 else torch.bfloat16



def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise capi.LavbError("lav_b200 kernels need CUDA tensors (there is no CPU fallback)")


def launches():
    """number of kernel launches issued through this module (bench.py reports it)."""
    return _COUNT[0]


_COUNT = [0]
PROFILE = None     # bench.py sets this to a list to collect (kind, work, start_event, end_event) per launch


def _prof_begin():
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(kind, work, e0):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        PROFILE.append((kind, work, e0, e1))


# ----------------------------------------------------------------------------- painting
def paint(points, sem, cams, mode, copy_cols=0, out=None, out_col0=None):
    """points (N,>=3) fp32; sem (ncam,C,H,W)-shaped tensor with ANY strides (NCHW or channels-last);
    cams (ncam,41) float32 numpy (K|lidar_to_world|world_to_cam).  See lavb_paint in include/lav_b200.h."""
    _need_cuda(points, sem)
    assert points.dtype == torch.float32 and sem.dtype == torch.float32 and points.dim() == 2
    assert points.stride(1) == 1
    ncam, c_in, h, w = sem.shape
    c_out = c_in if mode == 0 else c_in - 1
    n = points.shape[0]
    if out_col0 is None:
        out_col0 = copy_cols
    if out is None:
        out = torch.empty((n, out_col0 + c_out), dtype=torch.float32, device=points.device)
    assert out.stride(1) == 1
    cams = np.ascontiguousarray(cams, dtype=np.float32)
    assert cams.shape == (ncam, 41)
    s = sem.stride()
    check(lib().lavb_paint(_ptr(points), n, points.stride(0), _ptr(sem), ncam, c_in, h, w, s[0], s[1], s[2], s[3],
                           cams.ctypes.data_as(C.c_void_p), mode, _ptr(out), out.stride(0), out_col0, copy_cols, _stream()),
          "lavb_paint")
    _COUNT[0] += 1
    return out


def stack_sweep(src, R, dx, dy, time_idx, n_time, dst, roof_filter=False):
    """dst (n, src_cols+n_time) <- [src[:, :3] @ R + (dx,dy,0) | src[:,3:] | one_hot(time_idx)]"""
    _need_cuda(src, dst)
    assert src.is_contiguous() and dst.is_contiguous() and dst.shape == (src.shape[0], src.shape[1] + n_time)
    R = np.ascontiguousarray(R, dtype=np.float32)
    check(lib().lavb_stack_sweep(_ptr(src), src.shape[0], src.shape[1], R.ctypes.data_as(C.c_void_p), float(dx), float(dy),
                                 time_idx, n_time, int(roof_filter), _ptr(dst), _stream()), "lavb_stack_sweep")
    _COUNT[0] += 1
    return dst


def roof_filter(sweeps, pad_nan=False, out=None):
    """LAVAgent.preprocess (lav_agent.py:448-457) on the device, order preserving.  sweeps: (n, cols) or (F, n, cols) fp32
    contiguous -> (out like sweeps with the kept rows first, counts (F,) int32).  pad_nan fills rows past the count with NaN."""
    _need_cuda(sweeps)
    assert sweeps.dtype == torch.float32 and sweeps.is_contiguous() and sweeps.dim() in (2, 3)
    x = sweeps if sweeps.dim() == 3 else sweeps[None]
    f, n, cols = x.shape
    if out is None:
        out = torch.empty_like(x)
    assert out.is_contiguous() and out.shape == x.shape and out.data_ptr() != x.data_ptr()
    counts = torch.empty((f,), dtype=torch.int32, device=x.device)
    check(lib().lavb_roof_filter(_ptr(x), f, n, cols, n * cols, _ptr(out), n * cols, _ptr(counts), int(pad_nan), _stream()),
          "lavb_roof_filter")
    _COUNT[0] += 1
    return (out if sweeps.dim() == 3 else out[0]), counts


# ----------------------------------------------------------------------------- pillars
def _workspace(device, nbytes):
    """scratch for one call.  Deliberately NOT cached globally: under CUDA-graph capture the buffer must belong to the
    capturing graph's private pool (two pipelines replaying on different streams must never share scratch); the
    caching allocator makes the eager-mode cost negligible."""
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


def _clouds(starts, counts):
    b = len(counts)
    st = (C.c_longlong * b)(*[int(s) for s in starts])
    ct = (C.c_int * b)(*[int(c) for c in counts])
    return b, st, ct


def pillar_forward(pts, starts, counts, grid, w1, s1, t1, w2, s2, t2):
    """pts: 2-D fp32 row buffer (rows of >= D floats); cloud b = rows [starts[b], starts[b]+counts[b]).
    Returns the NHWC canvas (B, ny, nx, H2) fp32."""
    _need_cuda(pts, w1, w2)
    assert pts.dtype == torch.float32 and pts.dim() == 2 and pts.stride(1) == 1
    min_x, max_x, min_y, max_y, ppm, nx, ny = grid
    d = w1.shape[1] - 5
    b, st, ct = _clouds(starts, counts)
    canvas = torch.empty((b, ny, nx, w2.shape[0]), dtype=torch.float32, device=pts.device)
    ws = _workspace(pts.device, lib().lavb_pillar_workspace_bytes(b, nx, ny))
    e0 = _prof_begin()
    check(lib().lavb_pillar_forward(_ptr(pts), pts.stride(0), d, st, ct, b, min_x, max_x, min_y, max_y, ppm, nx, ny,
                                    _ptr(w1), _ptr(s1), _ptr(t1), w1.shape[0], _ptr(w2), _ptr(s2), _ptr(t2), w2.shape[0],
                                    _ptr(canvas), F32, _ptr(ws), _stream()), "lavb_pillar_forward")
    # algorithmic bytes: read P x D fp32 points once + write the canvas once
    _prof_end("pillar", float(sum(int(c) for c in counts)) * d * 4 + canvas.numel() * 4, e0)
    _COUNT[0] += 4
    return canvas


def pillar_decorate(pts, starts, counts, grid, d):
    """training stage 0: returns (feat (M,d+5) fp32, cell (M,) int32)."""
    _need_cuda(pts)
    min_x, max_x, min_y, max_y, ppm, nx, ny = grid
    b, st, ct = _clouds(starts, counts)
    ws = _workspace(pts.device, lib().lavb_pillar_workspace_bytes(b, nx, ny))
    total = int(sum(int(c) for c in counts))
    feat = torch.empty((total, d + 5), dtype=torch.float32, device=pts.device)
    cell = torch.empty((total,), dtype=torch.int32, device=pts.device)
    m = C.c_int(0)
    check(lib().lavb_pillar_decorate(_ptr(pts), pts.stride(0), d, st, ct, b, min_x, max_x, min_y, max_y, ppm, nx, ny,
                                     _ptr(feat), _ptr(cell), C.byref(m), _ptr(ws), _stream()), "lavb_pillar_decorate")
    _COUNT[0] += 4
    return feat[:m.value], cell[:m.value]


def pillar_scatter_max(h, cell, n_cells, want_argmax=True):
    _need_cuda(h, cell)
    h = h.contiguous()
    m, c = h.shape
    canvas = torch.empty((n_cells, c), dtype=torch.float32, device=h.device)
    arg = torch.empty((n_cells, c), dtype=torch.int32, device=h.device) if want_argmax else None
    check(lib().lavb_pillar_scatter_max(_ptr(h), _ptr(cell), m, c, n_cells, _ptr(canvas), _ptr(arg), _stream()),
          "lavb_pillar_scatter_max")
    _COUNT[0] += 2
    return canvas, arg


def pillar_scatter_max_bwd(gcanvas, arg, cell, m):
    gcanvas = gcanvas.contiguous()
    c = gcanvas.shape[-1]
    gh = torch.empty((m, c), dtype=torch.float32, device=gcanvas.device)
    check(lib().lavb_pillar_scatter_max_bwd(_ptr(gcanvas), _ptr(arg), _ptr(cell), m, c, _ptr(gh), _stream()),
          "lavb_pillar_scatter_max_bwd")
    _COUNT[0] += 1
    return gh


# ----------------------------------------------------------------------------- convolution
def conv_taps(x, cin, in_coff, out, cout, out_coff, hog, wog, in_s, out_s, out_o, taps, w, bias=None, scale=None, shift=None,
              res=None, res_coff=0, pre_relu=False, post_relu=False, sigmoid=False, umma=False, d2s_nout=0):
    """x, out, res: contiguous NHWC buffers (N,H,W,Ctot).  taps: list of (dy,dx).
    umma=False: CUDA-core kernel, w (ntaps,cin,cout_pad16) fp32.
    umma=True : tcgen05 kernel, x f16, w (ntaps,cout,cin) f16."""
    _need_cuda(x, out, w)
    assert x.is_contiguous() and out.is_contiguous() and w.is_contiguous()
    d = ConvDesc()
    d.inp, d.in_dtype = x.data_ptr(), _DT[x.dtype]
    d.n, d.hin, d.win, d.in_cstride = x.shape
    d.cin, d.in_coff = cin, in_coff
    d.out, d.out_dtype = out.data_ptr(), _DT[out.dtype]
    assert out.shape[0] == x.shape[0]
    _, d.hout, d.wout, d.out_cstride = out.shape
    d.cout, d.out_coff = cout, out_coff
    d.d2s_nout = d2s_nout
    if d2s_nout:
        assert umma and out.dtype == torch.float32 and out.shape[3] == d2s_nout and cout == 32 and 4 * d2s_nout <= 32
        d.out_cstride, d.out_coff = 32, 0        # (validated as a 32-column GEMM; addressing is done by the d2s epilogue)
    d.hog, d.wog = hog, wog
    d.in_sy, d.in_sx = in_s
    d.out_sy, d.out_sx = out_s
    d.out_oy, d.out_ox = out_o
    d.ntaps = len(taps)
    if umma:
        assert w.dtype == h16() and tuple(w.shape) == (len(taps), (cout + 31) // 32 * 32, cin) and x.dtype == h16()
    else:
        assert w.dtype == torch.float32 and tuple(w.shape) == (len(taps), cin, (cout + 15) // 16 * 16)
    for i, (dy, dx) in enumerate(taps):
        d.dy[i], d.dx[i] = dy, dx
    d.w = w.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    d.scale = scale.data_ptr() if scale is not None else None
    d.shift = shift.data_ptr() if shift is not None else None
    if res is not None:
        assert res.is_contiguous() and res.shape[:3] == out.shape[:3]
        d.res, d.res_dtype, d.res_cstride, d.res_coff = res.data_ptr(), _DT[res.dtype], res.shape[3], res_coff
    d.pre_relu, d.post_relu, d.sigmoid = int(pre_relu), int(post_relu), int(sigmoid)
    if umma:
        e0 = _prof_begin()
        check(lib().lavb_conv_umma(C.byref(d), _stream()), "lavb_conv_umma")
        _prof_end(f"umma:{cin}->{cout}x{len(taps)}taps@{hog}x{wog}", 2.0 * x.shape[0] * hog * wog * cout * cin * len(taps), e0)
    else:
        check(lib().lavb_conv_taps(C.byref(d), _stream()), "lavb_conv_taps")
    _COUNT[0] += 1
    return out


def pool2_affine_relu(x, c, in_coff, scale, shift, out, out_coff):
    _need_cuda(x, out)
    n, h, w, cs = x.shape
    check(lib().lavb_pool2_affine_relu(_ptr(x), _DT[x.dtype], n, h, w, c, cs, in_coff, _ptr(scale), _ptr(shift), _ptr(out),
                                       out.shape[3], out_coff, _stream()), "lavb_pool2_affine_relu")
    _COUNT[0] += 1
    return out


def rgb_normalize(rgb, out_dtype=torch.float32):
    """uint8 (N,H,W,3) or float (N,3,H,W) in 0..255 -> NHWC4 normalised ((x/255-.5)*2, 4th channel 0)."""
    _need_cuda(rgb)
    if rgb.dtype == torch.uint8:
        assert rgb.dim() == 4 and rgb.shape[3] == 3
        n, h, w, _ = rgb.shape
        u8 = 1
        rgb = rgb.contiguous()
    else:
        assert rgb.dim() == 4 and rgb.shape[1] == 3
        n, _, h, w = rgb.shape
        u8 = 0
        rgb = rgb.float().contiguous()
    out = torch.empty((n, h, w, 4), dtype=out_dtype, device=rgb.device)
    check(lib().lavb_rgb_normalize(_ptr(rgb), u8, n, h, w, _ptr(out), _DT[out_dtype], _stream()), "lavb_rgb_normalize")
    _COUNT[0] += 1
    return out


def convert(src, dtype):
    _need_cuda(src)
    src = src.contiguous()
    if src.dtype == dtype:
        return src
    dst = torch.empty(src.shape, dtype=dtype, device=src.device)
    check(lib().lavb_convert(_ptr(src), _DT[src.dtype], _ptr(dst), _DT[dtype], src.numel(), _stream()), "lavb_convert")
    _COUNT[0] += 1
    return dst


def crop_bilinear(feats_nhwc, frame_idx, theta, crop_size):
    """feats_nhwc (B,H,W,C) contiguous fp32/f16; frame_idx (K,) int32; theta (K,2,3) fp32 -> (K,crop,crop,C)."""
    _need_cuda(feats_nhwc, frame_idx, theta)
    assert feats_nhwc.is_contiguous()
    b, h, w, c = feats_nhwc.shape
    k = theta.shape[0]
    theta = theta.float().contiguous()
    frame_idx = frame_idx.to(torch.int32).contiguous()
    out = torch.empty((k, crop_size, crop_size, c), dtype=feats_nhwc.dtype, device=feats_nhwc.device)
    check(lib().lavb_crop_bilinear(_ptr(feats_nhwc), _DT[feats_nhwc.dtype], b, h, w, c, _ptr(frame_idx), _ptr(theta), k, crop_size,
                                   _ptr(out), _stream()), "lavb_crop_bilinear")
    _COUNT[0] += 1
    return out


def crop_bilinear_bwd(gout_nhwc, frame_idx, theta, feat_shape):
    """gout_nhwc (K,crop,crop,C) fp32 contiguous -> gradient of crop_bilinear w.r.t. the (B,H,W,C) fp32 feature map."""
    _need_cuda(gout_nhwc, frame_idx, theta)
    assert gout_nhwc.is_contiguous() and gout_nhwc.dtype == torch.float32
    b, h, w, c = feat_shape
    k, crop = gout_nhwc.shape[0], gout_nhwc.shape[1]
    gfeat = torch.empty((b, h, w, c), dtype=torch.float32, device=gout_nhwc.device)
    check(lib().lavb_crop_bilinear_bwd(_ptr(gout_nhwc), b, h, w, c, _ptr(frame_idx), _ptr(theta), k, crop, _ptr(gfeat), _stream()),
          "lavb_crop_bilinear_bwd")
    _COUNT[0] += 1
    return gfeat


class CropBilinear(torch.autograd.Function):
    """crop_bilinear with its hand-written backward (gradient to the feature map only: the crop poses are data)."""

    @staticmethod
    def forward(ctx, feats_nhwc, frame_idx, theta, crop_size):
        frame_idx = frame_idx.to(torch.int32).contiguous()
        theta = theta.detach().float().contiguous()
        ctx.save_for_backward(frame_idx, theta)
        ctx.feat_shape = tuple(feats_nhwc.shape)
        return crop_bilinear(feats_nhwc, frame_idx, theta, crop_size)

    @staticmethod
    def backward(ctx, gout):
        frame_idx, theta = ctx.saved_tensors
        return crop_bilinear_bwd(gout.contiguous(), frame_idx, theta, ctx.feat_shape), None, None, None


def deconv3x3s2_small(x, groups, cin_g, w, bias, n_outs, sigmoids):
    """x NHWC (N,H,W,Ctot); w fp32 (G,cin_g,9,4); bias (G,4) -> list of fp32 NHWC (N,2H,2W,n_out[g])."""
    _need_cuda(x, w, bias)
    assert x.is_contiguous() and w.is_contiguous() and bias.is_contiguous()
    n, h, wd, cs = x.shape
    outs = [torch.empty((n, 2 * h, 2 * wd, no), dtype=torch.float32, device=x.device) for no in n_outs]
    ptrs = (C.c_void_p * groups)(*[o.data_ptr() for o in outs])
    no = (C.c_int * groups)(*n_outs)
    sg = (C.c_int * groups)(*[int(s) for s in sigmoids])
    check(lib().lavb_deconv3x3s2_small(_ptr(x), _DT[x.dtype], n, h, wd, cs, groups, cin_g, _ptr(w), _ptr(bias), no, sg, ptrs,
                                       _stream()), "lavb_deconv3x3s2_small")
    _COUNT[0] += 1
    return outs


def paint_batched(points, sem, cams, mode, copy_cols, out):
    """points (F,N,>=3) fp32 contiguous; sem logical (F,ncam,C,H,W) any strides; out (F,N,copy_cols+c_out) contiguous."""
    _need_cuda(points, sem, out)
    assert points.is_contiguous() and out.is_contiguous() and points.dtype == torch.float32 and sem.dtype == torch.float32
    f, n, ps = points.shape
    _, ncam, c_in, h, w = sem.shape
    cams = np.ascontiguousarray(cams, dtype=np.float32)
    s = sem.stride()
    check(lib().lavb_paint_batched(_ptr(points), f, n, ps, n * ps, _ptr(sem), ncam, c_in, h, w, s[0], s[1], s[2], s[3], s[4],
                                   cams.ctypes.data_as(C.c_void_p), mode, _ptr(out), out.shape[2], n * out.shape[2], copy_cols,
                                   copy_cols, _stream()), "lavb_paint_batched")
    _COUNT[0] += 1
    return out


def pack_deconv2x2(weight, bias):
    """ConvTranspose2d(16, C, 2, stride=2) parameters (weight (16,C,2,2), bias (C,)) -> the 520-float table
    lavb_paint_deconv_batched reads: w[v%2][u%2][c_in][8] | bias[8]."""
    cin, c, kh, kw = weight.shape
    assert cin == 16 and kh == 2 and kw == 2 and c <= 8
    w = torch.zeros((2, 2, 16, 8), dtype=torch.float32, device=weight.device)
    w[:, :, :, :c] = weight.detach().float().permute(2, 3, 0, 1)
    b = torch.zeros((8,), dtype=torch.float32, device=weight.device)
    b[:c] = bias.detach().float()
    return torch.cat([w.reshape(-1), b]).contiguous()


def paint_deconv_batched(points, feat, n_classes, deconv, cams, copy_cols, out, image_hw):
    """points (F,N,>=3) fp32; feat NHWC (F*ncam, H/2, W/2, 16) fp32 / h16 = ERFNet decoder output before output_conv;
    deconv = pack_deconv2x2(...); out (F,N,copy_cols + n_classes-1)."""
    _need_cuda(points, feat, out, deconv)
    assert points.is_contiguous() and out.is_contiguous() and feat.is_contiguous() and points.dtype == torch.float32
    f, n, ps = points.shape
    h, w = image_hw
    cams = np.ascontiguousarray(cams, dtype=np.float32)
    ncam = cams.shape[0]
    assert feat.shape == (f * ncam, h // 2, w // 2, 16) and deconv.numel() == 520
    check(lib().lavb_paint_deconv_batched(_ptr(points), f, n, ps, n * ps, _ptr(feat), _DT[feat.dtype], ncam, n_classes, h, w,
                                          _ptr(deconv), cams.ctypes.data_as(C.c_void_p), _ptr(out), out.shape[2], n * out.shape[2],
                                          copy_cols, copy_cols, _stream()), "lavb_paint_deconv_batched")
    _COUNT[0] += 1
    return out


STACK_JOB_DTYPE = np.dtype([("src", np.uint64), ("dst", np.uint64), ("n", np.int32), ("time_idx", np.int32), ("R", np.float32, 9),
                            ("dx", np.float32), ("dy", np.float32), ("pad", np.int32)])
assert STACK_JOB_DTYPE.itemsize == 72


def stack_jobs(d_jobs, n_jobs, max_n, src_cols, n_time, roof_filter=False):
    """d_jobs: uint8 device tensor holding n_jobs STACK_JOB_DTYPE records."""
    _need_cuda(d_jobs)
    check(lib().lavb_stack_jobs(_ptr(d_jobs), n_jobs, max_n, src_cols, n_time, int(roof_filter), _stream()), "lavb_stack_jobs")
    _COUNT[0] += 1


def split_h16(x):
    """fp32 (..., C) contiguous -> f16 (..., 2C) = [hi | lo] error-free split (see lavb_split_h16)."""
    _need_cuda(x)
    assert x.is_contiguous() and x.dtype == torch.float32
    c = x.shape[-1]
    out = torch.empty((*x.shape[:-1], 2 * c), dtype=h16(), device=x.device)
    check(lib().lavb_split_h16(_ptr(x), _ptr(out), x.numel() // c, c, _stream()), "lavb_split_h16")
    _COUNT[0] += 1
    return out


PILLAR_ENCODER = "sorted"    # tensor-core encoders of the 16-bit pipeline, B200 @ 32 frames x 120 000 points:
#   "sorted": counting sort by canvas cell + persistent mma.sync encoder (lavb_pillar_forward_sorted)          23.3 us/frame
#   "tiled" : points binned by 8x16-cell canvas tile, one CTA per tile, tcgen05 MLP (lavb_pillar_forward_tiled) 29.2 us/frame —
#             fewer launches and 35 % less DRAM traffic (942 vs 1448 MB), but every tile is a serial chain of ~8 dependent steps
#             (load, centroid atomics, MMA round trips, pooling atomics, store) with 3 CTAs per SM: latency-bound (profiles/)


def pillar_forward_sorted(pts, starts, counts, grid, w1, s1, t1, w2, s2, t2, split_out=False, canvas16=False):
    """tensor-core pillar encoder of the 16-bit pipeline (tile-binned or sorted kernel, see PILLAR_ENCODER).  Returns the NHWC
    canvas: fp32 (B,ny,nx,H2); with split_out, h16 (B,ny,nx,2*H2) = [hi | lo]; with canvas16, h16 (B,ny,nx,H2) — what the 16-bit
    pipeline feeds the backbone."""
    _need_cuda(pts, w1, w2)
    assert pts.dtype == torch.float32 and pts.dim() == 2 and pts.stride(1) == 1
    min_x, max_x, min_y, max_y, ppm, nx, ny = grid
    d = w1.shape[1] - 5
    b, st, ct = _clouds(starts, counts)
    total = int(sum(int(c) for c in counts))
    h2 = w2.shape[0]
    assert not (split_out and canvas16)
    canvas = torch.empty((b, ny, nx, 2 * h2 if split_out else h2), dtype=h16() if (split_out or canvas16) else torch.float32, device=pts.device)
    tiled = PILLAR_ENCODER == "tiled"
    if tiled:      # every frame's records start at its exclusive point offset: size the record buffer for the clouds as given
        total = int(sum(int(c) for c in counts))
    ws = _workspace(pts.device, (lib().lavb_pillar_tiled_workspace_bytes if tiled else lib().lavb_pillar_sorted_workspace_bytes)(b, nx, ny, total))
    e0 = _prof_begin()
    fn, name = (lib().lavb_pillar_forward_tiled, "lavb_pillar_forward_tiled") if tiled else (lib().lavb_pillar_forward_sorted, "lavb_pillar_forward_sorted")
    check(fn(_ptr(pts), pts.stride(0), d, st, ct, b, min_x, max_x, min_y, max_y, ppm, nx, ny,
                                           _ptr(w1), _ptr(s1), _ptr(t1), w1.shape[0], _ptr(w2), _ptr(s2), _ptr(t2), h2,
                                           _ptr(canvas), 2 if canvas16 else int(split_out), _ptr(ws), _stream()), name)
    _prof_end("pillar", float(total) * d * 4 + float(b) * ny * nx * h2 * (2 if canvas16 else 4), e0)
    _COUNT[0] += 3 if tiled else 6
    return canvas


def det_peaks(center, box, ori, min_score=0.2, max_det=15):
    """center (LOGITS), box, ori: fp32 NHWC (B,H,W,2) contiguous -> packed (B,7,2*max_det) (see lavb_det_peaks)."""
    _need_cuda(center, box, ori)
    assert center.is_contiguous() and box.is_contiguous() and ori.is_contiguous() and center.dtype == torch.float32
    b, h, w, ncls = center.shape
    packed = torch.empty((b, 7, ncls * max_det), dtype=torch.float32, device=center.device)
    ws = _workspace(center.device, lib().lavb_det_peaks_workspace_bytes(b, ncls))
    check(lib().lavb_det_peaks(_ptr(center), _ptr(box), _ptr(ori), b, h, w, ncls, min_score, max_det, _ptr(packed), _ptr(ws), _stream()),
          "lavb_det_peaks")
    _COUNT[0] += 2
    return packed


def stem7x7s2_u8(img_u8, w_h16, bias, mean, std):
    """img_u8 (B, ncam, H, cam_w, 3) uint8 contiguous; w_h16 (64,160) packed by pack_stem_weights; bias (64,)
    -> f16 NHWC (B, H/2, ncam*cam_w/2, 64)."""
    _need_cuda(img_u8, w_h16, bias)
    assert img_u8.dtype == torch.uint8 and img_u8.is_contiguous() and img_u8.dim() == 5 and img_u8.shape[4] == 3
    assert w_h16.dtype == h16() and tuple(w_h16.shape) == (64, 160) and w_h16.is_contiguous()
    b, ncam, h, cw, _ = img_u8.shape
    out = torch.empty((b, (h - 1) // 2 + 1, (ncam * cw - 1) // 2 + 1, 64), dtype=h16(), device=img_u8.device)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    sd = (C.c_float * 3)(*[float(v) for v in std])
    check(lib().lavb_stem7x7s2_u8(_ptr(img_u8), b, ncam, h, cw, _ptr(w_h16), _ptr(bias), m, sd, _ptr(out), _stream()),
          "lavb_stem7x7s2_u8")
    _COUNT[0] += 1
    return out


def pack_stem_weights(w):
    """(64, 3, 7, 7) conv weights -> (64, 160) f16 in the stem kernel's K order k = ky*22 + kx*3 + c (zero elsewhere)."""
    wk = torch.zeros((64, 7, 22), dtype=torch.float32, device=w.device)
    wk[:, :, :21] = w.float().permute(0, 2, 3, 1).reshape(64, 7, 21)
    out = torch.zeros((64, 160), dtype=torch.float32, device=w.device)
    out[:, :154] = wk.reshape(64, 154)
    return out.to(h16()).contiguous()


def maxpool3x3s2_nhwc(x):
    """MaxPool2d(3, 2, 1) on a contiguous f16 NHWC tensor."""
    _need_cuda(x)
    assert x.dtype == h16() and x.is_contiguous() and x.dim() == 4 and x.shape[3] % 8 == 0
    n, h, w, c = x.shape
    out = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), dtype=h16(), device=x.device)
    check(lib().lavb_maxpool3x3s2_nhwc(_ptr(x), n, h, w, c, _ptr(out), _stream()), "lavb_maxpool3x3s2_nhwc")
    _COUNT[0] += 1
    return out


def conv_pair_umma(x, w1, bias1, w2, shift2, dil, res=None, post_relu=True):
    """fused pair: mid = relu(conv3x1_dil(x) + bias1); out = [relu](conv1x3_dil(mid) + shift2 [+ res]).  A BatchNorm affine after
    the second conv is folded by the caller: w2 <- w2 * s (per output channel), shift2 <- b2 * s + t.
    x / res: contiguous f16 NHWC (n, h, w, c), c in {64, 128}, w in {32, 64, 128}; w1 / w2: (3, c, c) f16 [tap][cout][cin];
    bias1 / shift2: fp32 (c,)."""
    from .capi import ConvPairDesc
    _need_cuda(x, w1, w2, bias1, shift2)
    n, h, w, c = x.shape
    assert x.dtype == h16() and x.is_contiguous() and w1.is_contiguous() and w2.is_contiguous()
    assert tuple(w1.shape) == (3, c, c) and tuple(w2.shape) == (3, c, c) and w1.dtype == w2.dtype == h16()
    assert bias1.dtype == shift2.dtype == torch.float32 and bias1.numel() == shift2.numel() == c
    out = torch.empty_like(x)
    d = ConvPairDesc()
    d.inp, d.out = x.data_ptr(), out.data_ptr()
    d.n, d.h, d.w, d.c, d.dil, d.post_relu = n, h, w, c, int(dil), int(post_relu)
    d.w1, d.bias1 = w1.data_ptr(), bias1.data_ptr()
    d.w2, d.shift2 = w2.data_ptr(), shift2.data_ptr()
    if res is not None:
        assert res.is_contiguous() and res.shape == x.shape and res.dtype == h16()
        d.res = res.data_ptr()
    e0 = _prof_begin()
    check(lib().lavb_conv_pair_umma(C.byref(d), _stream()), "lavb_conv_pair_umma")
    _prof_end(f"umma_pair:{c}x{h}x{w}", 2.0 * n * h * w * c * c * 6, e0)
    _COUNT[0] += 1
    return out


def erf_stem(rgb_u8, w27, scale, shift, out_dtype):
    """fused normalize + ERFNet initial block: rgb_u8 (N,H,W,3) uint8 -> NHWC (N,H/2,W/2,16).  w27 (27,16), scale/shift (16,)
    are HOST float32 numpy arrays (kernel parameters)."""
    _need_cuda(rgb_u8)
    assert rgb_u8.dtype == torch.uint8 and rgb_u8.is_contiguous() and rgb_u8.dim() == 4 and rgb_u8.shape[3] == 3
    n, h, w, _ = rgb_u8.shape
    out = torch.empty((n, h // 2, w // 2, 16), dtype=out_dtype, device=rgb_u8.device)
    a, b, c = (np.ascontiguousarray(t, dtype=np.float32) for t in (w27, scale, shift))
    assert a.shape == (27, 16) and b.shape == (16,) and c.shape == (16,)
    check(lib().lavb_erf_stem(_ptr(rgb_u8), n, h, w, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
                              _ptr(out), _DT[out_dtype], _stream()), "lavb_erf_stem")
    _COUNT[0] += 1
    return out


def erf_down16(x, w9, st):
    """fused DownsamplerBlock(16, 64): x h16 NHWC (n,h,w,16) -> (n,h/2,w/2,64) (lavb_erf_down16)."""
    _need_cuda(x, w9, st)
    assert x.dtype == h16() and x.is_contiguous() and x.dim() == 4 and x.shape[3] == 16
    assert w9.dtype == torch.float32 and tuple(w9.shape) == (9, 16, 48) and w9.is_contiguous()
    assert st.dtype == torch.float32 and tuple(st.shape) == (64, 2) and st.is_contiguous()
    n, h, w, _ = x.shape
    out = torch.empty((n, h // 2, w // 2, 64), dtype=x.dtype, device=x.device)
    check(lib().lavb_erf_down16(_ptr(x), _ptr(out), n, h, w, _ptr(w9), _ptr(st), _stream()), "lavb_erf_down16")
    _COUNT[0] += 1
    return out


def erf_nb16(x, w4, st, out=None):
    """fused non_bottleneck_1d(16, dilation 1) block: x h16 NHWC (n,h,w,16) -> same shape (lavb_erf_nb16)."""
    _need_cuda(x, w4, st)
    assert x.dtype == h16() and x.is_contiguous() and x.dim() == 4 and x.shape[3] == 16
    assert w4.dtype == torch.float32 and tuple(w4.shape) == (4, 3, 16, 16) and w4.is_contiguous()
    assert st.dtype == torch.float32 and tuple(st.shape) == (4, 16, 2) and st.is_contiguous()
    n, h, w, _ = x.shape
    if out is None:
        out = torch.empty_like(x)
    check(lib().lavb_erf_nb16(_ptr(x), _ptr(out), n, h, w, _ptr(w4), _ptr(st), _stream()), "lavb_erf_nb16")
    _COUNT[0] += 1
    return out


def cast_gru(embd, wih_t, whh_t, bih, bhh, wmlp, bmlp, steps):
    """the 6 cast branches in one launch (csrc/cast_gru.cu).  embd (N, 512) fp32; wih_t (ncmd, 512, 192), whh_t (ncmd, 64, 192) the
    TRANSPOSED GRU weights; bih / bhh (ncmd, 192); wmlp (ncmd, 2, 64); bmlp (ncmd, 2) -> (N, ncmd, steps, 2) fp32 cumulative waypoints."""
    _need_cuda(embd, wih_t, whh_t)
    n, ncmd = embd.shape[0], wih_t.shape[0]
    assert tuple(embd.shape) == (n, 512) and tuple(wih_t.shape) == (ncmd, 512, 192) and tuple(whh_t.shape) == (ncmd, 64, 192)
    assert tuple(bih.shape) == tuple(bhh.shape) == (ncmd, 192) and tuple(wmlp.shape) == (ncmd, 2, 64) and tuple(bmlp.shape) == (ncmd, 2)
    for t in (embd, wih_t, whh_t, bih, bhh, wmlp, bmlp):
        assert t.dtype == torch.float32 and t.is_contiguous()
    out = torch.empty((n, ncmd, steps, 2), dtype=torch.float32, device=embd.device)
    check(lib().lavb_cast_gru(_ptr(embd), n, _ptr(wih_t), _ptr(whh_t), _ptr(bih), _ptr(bhh), _ptr(wmlp), _ptr(bmlp), ncmd, steps,
                              _ptr(out), _stream()), "lavb_cast_gru")
    _COUNT[0] += 1
    return out


def gru_h512(u, h0, whh, wih, bih, bhh):
    """cluster-persistent GRU(4 -> 512) roll-out (csrc/gru_cluster.cu), fp32-class arithmetic.  u (N, T, 4), h0 (N, 512),
    whh (1536, 512), wih (1536, 4), bih / bhh (1536,), all fp32 -> out (N, T, 512) fp32 (the output sequence of
    nn.GRU(batch_first=True))."""
    _need_cuda(u, h0, whh)
    n, t, k = u.shape
    assert k == 4 and tuple(h0.shape) == (n, 512) and tuple(whh.shape) == (1536, 512)
    assert u.dtype == h0.dtype == whh.dtype == wih.dtype == bih.dtype == bhh.dtype == torch.float32
    u, h0 = u.contiguous(), h0.contiguous()
    assert whh.is_contiguous() and wih.is_contiguous()
    out = torch.empty((n, t, 512), dtype=torch.float32, device=u.device)
    check(lib().lavb_gru_h512(_ptr(u), _ptr(h0), _ptr(whh), _ptr(wih), _ptr(bih), _ptr(bhh), _ptr(out), n, t, _stream()),
          "lavb_gru_h512")
    _COUNT[0] += 1
    return out
