"""CPU oracle of the AlexNet hash-head forward (stage='val') -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference computes this path inside un-vendored TensorFlow-GPU 1.12.0 / cuDNN 7.2.1
(environment.yml:14-15,83-85); TensorFlow is not installable here and the reference has no test or golden vector
for any encoder value.  This file restates the graph from its call sites in plain PyTorch fp32
on the CPU, one function per reference step:

    main.py:144-148               normalize()            2*x/256 - 1        (de-quantisation noise injected explicitly or off)
    lib/util.py:12-21             preprocess_resize()    (x+1)*255.99/2, NCHW->NHWC, TF1 legacy bilinear to 256x256
    lib/architecture.py:215-249   ten_crop()             5 crops of the flipped image + 5 plain, minus the channel mean
    lib/architecture.py:253-359   conv/pool/LRN          HWIO weights, VALID/SAME, 2-group convs, LRN iff WGAN_SCALE == 0
    lib/architecture.py:363-389   fc6-8, tanh, crop mean (dropout masks injected explicitly or off; fc rows in (h,w,c) order)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

MEAN = (103.939, 116.779, 123.68)          # lib/architecture.py:247
CROP_OFFSETS = ((0, 0), (28, 28), (28, 0), (0, 28), (14, 14))  # (offset_height, offset_width), lib/architecture.py:216-241

SHAPES = {  # lib/architecture.py:253-382 (HWIO), SURVEY 8(a) row W
    "conv1": (11, 11, 3, 96), "conv2": (5, 5, 48, 256), "conv3": (3, 3, 256, 384), "conv4": (3, 3, 192, 384), "conv5": (3, 3, 192, 256),
    "fc6": (9216, 4096), "fc7": (4096, 4096),
}


def normalize(x_uint8: torch.Tensor, noise: torch.Tensor | None = None) -> torch.Tensor:
    x = 2 * x_uint8.to(torch.float32) / 256.0 - 1           # main.py:146
    if noise is not None:
        x = x + noise                                        # main.py:147 (U(0, 1/128)); None = deterministic mode
    return x


def tf1_resize_bilinear(img_nhwc: torch.Tensor, out_h: int = 256, out_w: int = 256) -> torch.Tensor:
    """tf.image.resize_bilinear of TF 1.12 with align_corners=False: src = dst * (in/out), no half-pixel centres."""
    n, in_h, in_w, c = img_nhwc.shape
    hs, ws = np.float32(in_h) / np.float32(out_h), np.float32(in_w) / np.float32(out_w)
    fy = (np.arange(out_h, dtype=np.float32) * hs).astype(np.float32)
    fx = (np.arange(out_w, dtype=np.float32) * ws).astype(np.float32)
    y0, x0 = np.floor(fy).astype(np.int64), np.floor(fx).astype(np.int64)
    y1, x1 = np.minimum(np.ceil(fy).astype(np.int64), in_h - 1), np.minimum(np.ceil(fx).astype(np.int64), in_w - 1)
    ly = torch.from_numpy(fy - y0.astype(np.float32)).view(1, out_h, 1, 1)
    lx = torch.from_numpy(fx - x0.astype(np.float32)).view(1, 1, out_w, 1)
    y0, y1, x0, x1 = (torch.from_numpy(a) for a in (y0, y1, x0, x1))
    tl, tr = img_nhwc[:, y0][:, :, x0], img_nhwc[:, y0][:, :, x1]
    bl, br = img_nhwc[:, y1][:, :, x0], img_nhwc[:, y1][:, :, x1]
    top = tl + (tr - tl) * lx
    bot = bl + (br - bl) * lx
    return top + (bot - top) * ly


def preprocess_resize(x: torch.Tensor, wh: int) -> torch.Tensor:
    img = (x + 1.0) * 255.99 / 2                              # lib/util.py:13
    img = img.reshape(-1, 3, wh, wh).permute(0, 2, 3, 1)      # lib/util.py:15-18
    return tf1_resize_bilinear(img.contiguous())              # lib/util.py:19


def ten_crop(img256: torch.Tensor) -> torch.Tensor:
    flipped = torch.flip(img256, dims=[2])                    # tf.image.flip_left_right
    crops = [flipped[:, oy:oy + 227, ox:ox + 227, :] for oy, ox in CROP_OFFSETS]
    crops += [img256[:, oy:oy + 227, ox:ox + 227, :] for oy, ox in CROP_OFFSETS]
    out = torch.cat(crops, 0)                                 # lib/architecture.py:242-244
    return out - torch.tensor(MEAN, dtype=torch.float32).view(1, 1, 1, 3)  # :247-249


def _conv(x_nhwc, w_hwio, b, stride, pad, groups):
    w = torch.as_tensor(w_hwio).permute(3, 2, 0, 1).contiguous()   # HWIO -> OIHW; torch groups == TF split/concat on channels
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w, torch.as_tensor(b), stride=stride, padding=pad, groups=groups)
    return F.relu(y).permute(0, 2, 3, 1).contiguous()


def _pool(x_nhwc):
    return F.max_pool2d(x_nhwc.permute(0, 3, 1, 2), 3, 2).permute(0, 2, 3, 1).contiguous()


def _lrn(x_nhwc):
    # tf.nn.local_response_normalization(depth_radius=2, alpha=2e-05, beta=0.75, bias=1.0): alpha is NOT divided by the window,
    # torch divides by size=5 -> alpha=1e-4
    return F.local_response_norm(x_nhwc.permute(0, 3, 1, 2), size=5, alpha=1e-4, beta=0.75, k=1.0).permute(0, 2, 3, 1).contiguous()


def _dropout(x, keep):
    # tf.nn.dropout(x, 0.5), lib/architecture.py:369,377: kept activations are scaled by 1 / keep_prob = 2; keep=None: off
    return x if keep is None else x * torch.as_tensor(np.asarray(keep), dtype=torch.float32) * 2.0


def encode(images_uint8, weights: dict, wh: int, lrn: bool = True, noise=None, return_pre_tanh: bool = False, keep6=None, keep7=None):
    """images_uint8: [B, 3*wh*wh] or [B, 3, wh, wh] (RGB planes).  weights: name -> array with the reference's names
    ('discriminator.conv1.weights', ..., 'discriminator.ACGANOutput.W').  Returns float32 [B, HASH_DIM]."""
    with torch.no_grad():
        x = torch.as_tensor(np.asarray(images_uint8)).reshape(len(images_uint8), -1)
        B = x.shape[0]
        g = lambda k: torch.as_tensor(np.asarray(weights[k], dtype=np.float32))
        x = normalize(x, None if noise is None else torch.as_tensor(np.asarray(noise, dtype=np.float32)).reshape(B, -1))
        x = preprocess_resize(x, wh)
        x = ten_crop(x)
        x = _conv(x, g("discriminator.conv1.weights"), g("discriminator.conv1.biases"), 4, 0, 1)   # :253-258
        x = _pool(x)                                                                               # :261-265
        if lrn:
            x = _lrn(x)                                                                            # :268-271
        x = _conv(x, g("discriminator.conv2.weights"), g("discriminator.conv2.biases"), 1, 2, 2)   # :275-288
        x = _pool(x)
        if lrn:
            x = _lrn(x)
        x = _conv(x, g("discriminator.conv3.weights"), g("discriminator.conv3.biases"), 1, 1, 1)   # :313-318
        x = _conv(x, g("discriminator.conv4.weights"), g("discriminator.conv4.biases"), 1, 1, 2)   # :322-335
        x = _conv(x, g("discriminator.conv5.weights"), g("discriminator.conv5.biases"), 1, 1, 2)   # :339-351
        x = _pool(x)                                                                               # :354-359
        x = x.reshape(x.shape[0], -1)                                                              # (h, w, c) flatten, :367
        x = _dropout(F.relu(x @ g("discriminator.fc6.weights") + g("discriminator.fc6.biases")), keep6)   # :368-369
        x = _dropout(F.relu(x @ g("discriminator.fc7.weights") + g("discriminator.fc7.biases")), keep7)   # :376-377
        fc8 = x @ g("discriminator.ACGANOutput.W") + g("discriminator.ACGANOutput.b")               # :381-382, lib/ops.py:287-302
        if return_pre_tanh:
            return fc8.reshape(10, B, -1).numpy()
        out = torch.tanh(fc8).reshape(10, B, -1).mean(0)                                            # :386-389
        return out.numpy()
