"""Drop-in for the reference's lib/core/inference.py (get_max_preds / get_final_preds) with the arg-max on the
device: only [B,J] indices and maxima cross PCIe instead of the full heat-maps (reference: D2H of
B*J*h*w floats per batch + numpy, inference.py:18-46)."""
import math

import numpy as np
import torch


def _device_argmax(hm):
    from fpd_b200 import ops
    idx, maxvals = ops.argmax_nchw(hm.contiguous().float())
    return idx, maxvals


def preds_from_argmax(idx, maxvals, width):
    """idx [B,J] flat arg-max, maxvals [B,J] -> (preds [B,J,2] float32, maxvals [B,J,1]) exactly as
    inference.py:36-46: x = idx % W, y = floor(idx / W), zeroed where max <= 0."""
    idx = np.asarray(idx).astype(np.float32)
    maxvals = np.asarray(maxvals, dtype=np.float32)[..., None]
    preds = np.stack([idx % width, np.floor(idx / width)], axis=-1).astype(np.float32)
    preds *= (maxvals > 0.0).astype(np.float32)
    return preds, maxvals


def get_max_preds(batch_heatmaps):
    """Accepts a CUDA tensor [B,J,h,w] (arg-max on the device) or, like the reference, a numpy array."""
    if isinstance(batch_heatmaps, np.ndarray):
        assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
        batch_heatmaps = torch.from_numpy(np.ascontiguousarray(batch_heatmaps)).cuda()
    idx, maxvals = _device_argmax(batch_heatmaps)
    return preds_from_argmax(idx.cpu().numpy(), maxvals.cpu().numpy(), batch_heatmaps.shape[3])


def _inverse_affine(center, scale, output_size):
    """The 2x3 matrix the reference obtains from get_affine_transform(center, scale, 0, output_size, inv=1)
    (lib/utils/transforms.py:57-89): three corresponding points -- centre, centre shifted up by half the box width, and
    the perpendicular third point -- held in float32 exactly like the reference's `src` / `dst` arrays, then the 3-point
    affine solved in float64 (what cv2.getAffineTransform does). Only scale[0] enters (the reference derives both axes
    from the box width)."""
    c = np.asarray(center, np.float32)
    src_w = np.asarray(scale, np.float32)[0] * np.float32(200.0)
    dst_w, dst_h = output_size[0], output_size[1]
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    src[0] = c
    src[1] = c + np.array([0.0, float(src_w * np.float32(-0.5))])
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0, dst_w * -0.5], np.float32)
    for p in (src, dst):
        d = p[0] - p[1]
        p[2] = p[1] + np.array([-d[1], d[0]], np.float32)
    a = np.concatenate([dst.astype(np.float64), np.ones((3, 1))], 1)
    return np.linalg.solve(a, src.astype(np.float64)).T          # maps heat-map (dst) coordinates to image (src)


def transform_preds(coords, center, scale, output_size):
    """lib/utils/transforms.py:49-54 for the whole [J,2] array at once (rot = 0 is the only call the hot path makes)."""
    coords = np.asarray(coords, dtype=np.float64)
    t = _inverse_affine(center, scale, output_size)
    out = np.zeros(coords.shape)
    out[:, 0:2] = coords[:, 0:2] @ t[:, :2].T + t[:, 2]
    return out


def get_final_preds(config, batch_heatmaps, center, scale):
    """inference.py:49-79: arg-max, optional quarter-pixel nudge toward the higher neighbour, map back to
    image coordinates. batch_heatmaps: CUDA tensor or numpy [B,J,h,w]."""
    if isinstance(batch_heatmaps, np.ndarray):
        batch_heatmaps = torch.from_numpy(np.ascontiguousarray(batch_heatmaps)).cuda()
    hm = batch_heatmaps.contiguous().float()
    B, J, H, W = hm.shape
    idx, maxvals = _device_argmax(hm)
    coords, maxvals_np = preds_from_argmax(idx.cpu().numpy(), maxvals.cpu().numpy(), W)
    if config.TEST.POST_PROCESS:
        # gather the four neighbours of every arg-max on the device: 4*B*J floats cross PCIe, not B*J*H*W
        px = torch.floor(torch.from_numpy(coords[..., 0]).cuda() + 0.5).long()
        py = torch.floor(torch.from_numpy(coords[..., 1]).cuda() + 0.5).long()
        ok = (px > 1) & (px < W - 1) & (py > 1) & (py < H - 1)
        pxc, pyc = px.clamp(1, W - 2), py.clamp(1, H - 2)
        flat = hm.reshape(B, J, H * W)

        def at(yy, xx):
            return flat.gather(2, (yy * W + xx).unsqueeze(-1)).squeeze(-1)
        dx = torch.sign(at(pyc, pxc + 1) - at(pyc, pxc - 1)) * ok
        dy = torch.sign(at(pyc + 1, pxc) - at(pyc - 1, pxc)) * ok
        coords[..., 0] += dx.cpu().numpy() * .25
        coords[..., 1] += dy.cpu().numpy() * .25
    preds = coords.copy()
    for i in range(B):
        preds[i] = transform_preds(coords[i], center[i], scale[i], [W, H])
    return preds, maxvals_np
