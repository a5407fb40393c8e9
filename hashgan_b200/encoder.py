"""AlexNet hash-head encoder (stage='val') -- host mirror of the reference's encode operator.

Reference: the per-batch operator `session.run(model.disc_real_acgan, {labeled_real_data_holder: image, ...})`
(main.py:154-155) = Model.normalize (main.py:144-148) -> discriminator(stage='val') = alexnet_discriminator
(lib/architecture.py:196-392).  Weights keep the reference's registry names (lib/params.py:11-35):
    discriminator.conv{1..5}.{weights,biases}   HWIO, shapes of lib/architecture.py:253-341
    discriminator.fc{6,7}.{weights,biases}
    discriminator.ACGANOutput.{W,b}             lib/ops.py:264-266,296-301 (fc8, [4096, HASH_DIM])
The numeric work is libhashgan_b200's hg_alexnet_encode (conv1-5 CUDA kernels, fc6-8 on tcgen05 tensor cores).
Deterministic mode only (EVAL.DETERMINISTIC): no de-quantisation noise, no eval-time dropout.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _native

__all__ = ["AlexNetWeights", "AlexNetHashEncoder", "CONV_SHAPES", "WEIGHT_NAMES", "stochastic_draws"]

_M64 = (1 << 64) - 1
_STREAM_NOISE, _STREAM_DROP6, _STREAM_DROP7 = 0xD1B54A32D192ED03, 0xA24BAED4963EE407, 0x9FB21C651E98DF25


def _mix(seed: int, idx: np.ndarray) -> np.ndarray:
    """hg_mix of csrc/encoder.cu (splitmix64 finaliser of seed + golden * (idx + 1)) on uint64 arrays."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _M64) + np.uint64(0x9E3779B97F4A7C15) * (idx.astype(np.uint64) + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def stochastic_draws(seed: int, n: int, wh: int):
    """The random draws hg_alexnet_encode_stochastic makes for a batch of n images with `seed`:
    (noise [n, 3*wh*wh] float32 = main.py:147, keep6 [10n, 4096] bool, keep7 [10n, 4096] bool = architecture.py:369,377)."""
    noise = (_mix(seed ^ _STREAM_NOISE, np.arange(n * 3 * wh * wh)) >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24) * np.float32(1 / 128)
    idx = np.arange(10 * n * 4096)
    keep6 = (_mix(seed ^ _STREAM_DROP6, idx) >> np.uint64(63)).astype(bool).reshape(10 * n, 4096)
    keep7 = (_mix(seed ^ _STREAM_DROP7, idx) >> np.uint64(63)).astype(bool).reshape(10 * n, 4096)
    return noise.reshape(n, -1), keep6, keep7

CONV_SHAPES = {"conv1": (11, 11, 3, 96), "conv2": (5, 5, 48, 256), "conv3": (3, 3, 256, 384), "conv4": (3, 3, 192, 384),
               "conv5": (3, 3, 192, 256)}
FC_SHAPES = {"fc6": (9216, 4096), "fc7": (4096, 4096)}


def WEIGHT_NAMES(hash_dim: int) -> Dict[str, tuple]:
    names = {}
    for k, s in CONV_SHAPES.items():
        names[f"discriminator.{k}.weights"] = s
        names[f"discriminator.{k}.biases"] = (s[3],)
    for k, s in FC_SHAPES.items():
        names[f"discriminator.{k}.weights"] = s
        names[f"discriminator.{k}.biases"] = (s[1],)
    names["discriminator.ACGANOutput.W"] = (4096, hash_dim)
    names["discriminator.ACGANOutput.b"] = (hash_dim,)
    return names


class AlexNetWeights:
    """name -> float32 array, validated against the reference's shapes."""

    def __init__(self, tensors: Dict[str, np.ndarray], hash_dim: int):
        want = WEIGHT_NAMES(hash_dim)
        self.hash_dim = hash_dim
        self.tensors = {}
        for name, shape in want.items():
            if name not in tensors:
                raise KeyError(f"missing weight {name}")
            a = np.ascontiguousarray(np.asarray(tensors[name], dtype=np.float32))
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {shape}, got {tuple(a.shape)}")
            self.tensors[name] = a

    @classmethod
    def synthetic(cls, hash_dim: int, seed: int = 0) -> "AlexNetWeights":
        """Seeded He-normal conv/fc weights, Glorot-uniform fc8 like lib/ops.py:213-218, small biases."""
        rng = np.random.default_rng(seed)
        t = {}
        for name, shape in WEIGHT_NAMES(hash_dim).items():
            if name.endswith(".W"):
                lim = np.sqrt(2.0 / (shape[0] + shape[1])) * np.sqrt(3.0)
                t[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
            elif len(shape) > 1:
                fan_in = int(np.prod(shape[:-1]))
                gain = np.sqrt(2.0 / fan_in)
                if name == "discriminator.conv1.weights":
                    gain /= 200.0  # mean-subtracted pixels are O(100): bring activations (and fc8) to O(1) so tanh is not saturated
                t[name] = (rng.standard_normal(size=shape, dtype=np.float32) * gain).astype(np.float32)
            else:
                t[name] = (rng.standard_normal(size=shape, dtype=np.float32) * 0.01).astype(np.float32)
        return cls(t, hash_dim)

    @classmethod
    def from_alexnet_npy(cls, path: str, hash_dim: int, seed: int = 0) -> "AlexNetWeights":
        """The reference's ImageNet initialisation: a pickled dict net_data[layer][0|1] (lib/architecture.py:199);
        fc8 (ACGANOutput) is not in that file and gets the reference's Glorot-uniform init (lib/ops.py:213-218)."""
        net = dict(np.load(path, encoding="latin1", allow_pickle=True).item())
        t = cls.synthetic(hash_dim, seed).tensors
        for layer in ("conv1", "conv2", "conv3", "conv4", "conv5", "fc6", "fc7"):
            t[f"discriminator.{layer}.weights"] = np.asarray(net[layer][0], dtype=np.float32)
            t[f"discriminator.{layer}.biases"] = np.asarray(net[layer][1], dtype=np.float32)
        t["discriminator.ACGANOutput.b"] = np.zeros((hash_dim,), dtype=np.float32)
        return cls(t, hash_dim)


    def override_from_tf_checkpoint(self, prefix: str, verify: bool = True, allow_partial: bool = False) -> list:
        """The reference's restore order (main.py:187-195): variables first take their .npy / initial values
        (lib/architecture.py:199), then `Saver.restore(session, D_PRETRAINED_MODEL_PATH)` overwrites every discriminator
        variable.  tf.train.Saver.restore raises NotFoundError when the checkpoint lacks a variable of the graph, so a
        checkpoint that misses (or renames) any of WEIGHT_NAMES(hash_dim) raises KeyError here as well -- it must not be
        evaluated with the initial values of that tensor.  `allow_partial=True` (tests, fine-tuning experiments) restores the
        intersection instead.  Returns the names that were overridden; a checkpoint tensor whose shape does not match the
        graph raises, as the restore op does."""
        from . import tf_checkpoint

        have = tf_checkpoint.list_variables(prefix)
        missing = [n for n in self.tensors if n not in have]
        if missing and not allow_partial:
            raise KeyError(f"checkpoint {prefix} lacks {len(missing)} variable(s) of the evaluation graph: {missing} "
                           "(tf.train.Saver.restore fails the same way, main.py:193); pass allow_partial=True to restore the rest")
        names = [n for n in self.tensors if n in have]
        for name, a in tf_checkpoint.read_checkpoint(prefix, names, verify=verify).items():
            if tuple(a.shape) != tuple(self.tensors[name].shape):
                raise ValueError(f"{name}: checkpoint shape {tuple(a.shape)} != graph shape {tuple(self.tensors[name].shape)}")
            self.tensors[name] = np.ascontiguousarray(a, dtype=np.float32)
        return names


class AlexNetHashEncoder:
    """images (uint8, [B, 3*wh*wh] as the loader yields them, lib/dataloader.py:110-113) -> CUDA float32 [B, HASH_DIM]."""

    def __init__(self, weights: AlexNetWeights, *, lrn: bool = True, device=None, conv_tf32: bool = False, deterministic: bool = True,
                 seed: int = 0, conv: Optional[str] = None, fused_stage1: bool = True):
        import torch

        if not torch.cuda.is_available():
            raise _native.NativeLibraryError("hashgan_b200 needs a CUDA device (sm_100a); there is no CPU fallback for the encoder")
        self.torch = torch
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.hash_dim = weights.hash_dim
        self.lrn = lrn
        # conv1-5: "tf32x3" (default) = implicit GEMM on tcgen05 with error-compensated TF32 (fp32-grade accuracy, 4x faster
        # than the CUDA cores); "fp32" = CUDA cores; "tf32" = plain TF32 operands (fastest, ~1e-3 relative error).
        # conv_tf32=True is the older spelling of conv="tf32".
        self.conv = conv if conv is not None else ("tf32" if conv_tf32 else "tf32x3")
        if self.conv not in ("fp32", "tf32", "tf32x3"):
            raise ValueError("conv must be 'fp32', 'tf32' or 'tf32x3'")
        conv_tf32 = self.conv != "fp32"
        self.conv_tf32 = conv_tf32
        # fused_stage1: normalise + resize + 10-crop + mean + conv1 + ReLU + pool1 (+LRN) as one kernel on effective filters
        # (csrc/encoder_stage1.cu) for 32 x 32 and 64 x 64 images; False = the separate kernels (any image size)
        self.fused_stage1 = bool(fused_stage1)
        self._fused = {}  # wh -> packed effective conv1 filters
        # deterministic=False: the reference's stochastic eval graph (de-quantisation noise main.py:147, dropout at eval
        # architecture.py:369,377); every encode() call draws with a fresh seed derived from `seed` and the call count
        self.deterministic = deterministic
        self.seed = int(seed)
        self.calls = 0
        self.last_seed = 0
        self.timing = False  # True: per-stage CUDA events in the next calls (hg_alexnet_phase_ms; bench.py)
        self.lib = _native.lib()
        dev = self.device
        t = {k: torch.from_numpy(v).to(dev) for k, v in weights.tensors.items()}
        self._keep = t
        st = _native.AlexNetWeightsStruct()
        for i in range(5):
            st.conv_w[i] = t[f"discriminator.conv{i + 1}.weights"].data_ptr()
            st.conv_b[i] = t[f"discriminator.conv{i + 1}.biases"].data_ptr()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            self._wt = {}
            for key, name in (("fc6", "discriminator.fc6.weights"), ("fc7", "discriminator.fc7.weights"), ("fc8", "discriminator.ACGANOutput.W")):
                src = t[name]
                dst = torch.empty((src.shape[1], src.shape[0]), dtype=torch.float32, device=dev)
                _native.check(self.lib.hg_transpose_f32(src.data_ptr(), src.shape[0], src.shape[1], dst.data_ptr(), stream))
                self._wt[key] = dst
            if conv_tf32:
                for i, (name, shp) in enumerate(CONV_SHAPES.items()):
                    kh, kw, cg, cout = shp
                    groups = 1 if name in ("conv1", "conv3") else 2
                    kpad = ((kh * kw * ((cg + 3) // 4 * 4) + 31) // 32) * 32
                    dst = torch.empty((2 * cout * kpad,), dtype=torch.float32, device=dev)  # [hi | lo], hg_conv_weight_pack
                    _native.check(self.lib.hg_conv_weight_pack(t[f"discriminator.{name}.weights"].data_ptr(), kh, kw, cg, cout, groups,
                                                               dst.data_ptr(), stream))
                    self._wt[name] = dst
                    st.conv_wt[i] = dst.data_ptr()
                if self.conv == "tf32x3":  # dense layers error-compensated too: [K, N] is HWIO with a 1x1 window
                    for i, name in enumerate(("discriminator.fc6.weights", "discriminator.fc7.weights", "discriminator.ACGANOutput.W")):
                        src = t[name]
                        k_in, n_out = int(src.shape[0]), int(src.shape[1])
                        dst = torch.empty((2 * n_out * k_in,), dtype=torch.float32, device=dev)
                        _native.check(self.lib.hg_conv_weight_pack(src.data_ptr(), 1, 1, k_in, n_out, 1, dst.data_ptr(), stream))
                        self._wt[f"fc3_{i}"] = dst
                        st.fc_wt3[i] = dst.data_ptr()
            torch.cuda.synchronize(dev)
        st.fc6_wt, st.fc6_b = self._wt["fc6"].data_ptr(), t["discriminator.fc6.biases"].data_ptr()
        st.fc7_wt, st.fc7_b = self._wt["fc7"].data_ptr(), t["discriminator.fc7.biases"].data_ptr()
        st.fc8_wt, st.fc8_b = self._wt["fc8"].data_ptr(), t["discriminator.ACGANOutput.b"].data_ptr()
        self._struct = st
        self._ws = None

    def encode(self, images, wh: Optional[int] = None):
        torch = self.torch
        if isinstance(images, torch.Tensor):
            x = images
        else:
            a = np.asarray(images)
            if a.dtype != np.uint8:
                if a.size and (a.min() < 0 or a.max() > 255):
                    raise ValueError("image values must be 0..255 (the loader yields uint8-valued pixels)")
                a = a.astype(np.uint8)
            x = torch.from_numpy(np.ascontiguousarray(a))
        n = int(x.shape[0])
        x = x.reshape(n, -1)
        if wh is None:
            wh = int(round((x.shape[1] / 3) ** 0.5))
        if 3 * wh * wh != x.shape[1]:
            raise ValueError(f"each image must have 3*wh*wh values, got {x.shape[1]}")
        if x.dtype != torch.uint8:
            x = x.to(torch.uint8)
        dev = self.device
        with torch.cuda.device(dev):
            x = x.to(dev, non_blocking=True).contiguous()
            out = torch.empty((n, self.hash_dim), dtype=torch.float32, device=dev)
            flags = (_native.ENC_LRN if self.lrn else 0) | {"fp32": 0, "tf32": _native.ENC_CONV_TF32, "tf32x3": _native.ENC_CONV_TF32X3}[self.conv]
            if self.timing:
                flags |= _native.ENC_TIMING
            if self.fused_stage1 and self.lib.hg_conv1_fused_floats(wh) > 0:
                if wh not in self._fused:
                    buf = torch.empty((self.lib.hg_conv1_fused_floats(wh),), dtype=torch.float32, device=dev)
                    _native.check(self.lib.hg_conv1_fused_pack(self._keep["discriminator.conv1.weights"].data_ptr(), wh, buf.data_ptr(),
                                                               torch.cuda.current_stream(dev).cuda_stream))
                    self._fused[wh] = buf
                self._struct.conv1_fused = self._fused[wh].data_ptr()
                self._struct.conv1_fused_wh = wh
                flags |= _native.ENC_FUSED_STAGE1
            need = self.lib.hg_alexnet_workspace_bytes(n, flags)
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty((need,), dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            if self.deterministic:
                _native.check(self.lib.hg_alexnet_encode(x.data_ptr(), n, wh, C.byref(self._struct), self.hash_dim, flags, out.data_ptr(),
                                                         self._ws.data_ptr(), self._ws.numel(), stream))
            else:
                self.last_seed = int(_mix(self.seed ^ 0x5851F42D4C957F2D, np.asarray([self.calls]))[0]) or 1
                self.calls += 1
                _native.check(self.lib.hg_alexnet_encode_stochastic(x.data_ptr(), n, wh, C.byref(self._struct), self.hash_dim, flags,
                                                                    out.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                                                    C.c_uint64(self.last_seed), stream))
            x.record_stream(torch.cuda.current_stream(dev))
        return out

    __call__ = encode
