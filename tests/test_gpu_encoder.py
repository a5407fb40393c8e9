"""Encoder-side kernels on a real B200: the tcgen05 (TF32) dense layer against a torch fp32 reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gemm(a, bt, bias, relu):
    import torch
    from hashgan_b200 import _native

    lib = _native.lib()
    M, K = a.shape
    N = bt.shape[0]
    c = torch.full((M, N), float("nan"), dtype=torch.float32, device=a.device)
    stream = torch.cuda.current_stream().cuda_stream
    _native.check(lib.hg_gemm_tf32(a.data_ptr(), a.stride(0), bt.data_ptr(), bt.stride(0), bias.data_ptr() if bias is not None else None,
                                   c.data_ptr(), c.stride(0), M, N, K, 1 if relu else 0, stream))
    torch.cuda.synchronize()
    return c


@pytest.mark.parametrize("M,N,K,relu", [(128, 128, 32, False), (128, 128, 256, False), (256, 384, 512, True), (1280, 4096, 1024, True),
                                        (1280, 64, 4096, False), (100, 48, 4096, False), (333, 200, 96, True), (1280, 4096, 9216, True)])
def test_gemm_tf32_matches_fp32_reference(M, N, K, relu):
    import torch

    torch.manual_seed(M * 7 + N * 3 + K)
    dev = torch.device("cuda:0")
    a = torch.randn(M, K, device=dev)
    bt = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    got = _gemm(a, bt, bias, relu)
    torch.backends.cuda.matmul.allow_tf32 = False
    want = a.double() @ bt.double().t() + bias.double()
    if relu:
        want = want.clamp_min(0)
    assert not torch.isnan(got).any()
    err = (got.double() - want).abs().max().item()
    scale = want.abs().max().item()
    # TF32 keeps 10 mantissa bits: |err| ~ 2^-11 * sqrt(K) * |a||b| per term; allow 2e-3 of the output scale
    assert err <= 2e-3 * max(scale, 1.0), (err, scale)
    # exact integers survive TF32: a stricter structural check (small integer operands are exact in tf32)
    ai = torch.randint(-3, 4, (M, K), device=dev).float()
    bi = torch.randint(-3, 4, (N, K), device=dev).float()
    goti = _gemm(ai, bi, None, False)
    wanti = (ai.double() @ bi.double().t()).float()
    assert torch.equal(goti, wanti)


def test_alexnet_encoder_matches_fp32_oracle():
    """configs[2] shape (CIFAR 32x32 -> 10 crops -> conv1-5 -> fc6-8 -> tanh -> crop mean) on synthetic weights/images:
    outputs within TF32 tolerance of the PyTorch fp32 restatement, code bits identical outside a small margin."""
    import torch
    from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights
    from oracle import alexnet_oracle

    for hash_dim, lrn, n in ((64, True, 12), (48, False, 5)):
        w = AlexNetWeights.synthetic(hash_dim, seed=3)
        rng = np.random.default_rng(5)
        img = rng.integers(0, 256, (n, 3 * 32 * 32), dtype=np.uint8)
        want = alexnet_oracle.encode(img, w.tensors, 32, lrn=lrn)
        enc = AlexNetHashEncoder(w, lrn=lrn, device="cuda:0")
        got = enc(img).cpu().numpy()
        assert got.shape == want.shape == (n, hash_dim)
        err = np.abs(got - want).max()
        assert err <= 5e-3, err  # TF32 fc layers (10-bit mantissa) vs fp32; tanh compresses the error
        margin = np.abs(want) > 2e-2
        assert np.array_equal(got[margin] > 0, want[margin] > 0)
        assert ((got > 0) == (want > 0)).mean() >= 0.995
        assert 0.2 < np.abs(want).mean() < 0.999  # the test exercises the unsaturated range of tanh


def test_encoder_at_the_configured_batch_of_128():
    """cifar_evaluation.yaml runs batches of 128 images (1280 crops) -- a different grid / tile regime from
    the small batches of the other tests.  Default (error-compensated tensor-core) path against the fp32 oracle, every image."""
    from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights
    from oracle import alexnet_oracle

    w = AlexNetWeights.synthetic(64, seed=21)
    img = np.random.default_rng(8).integers(0, 256, (128, 3 * 32 * 32), dtype=np.uint8)
    want = np.concatenate([alexnet_oracle.encode(img[i:i + 32], w.tensors, 32, lrn=True) for i in range(0, 128, 32)])
    enc = AlexNetHashEncoder(w, lrn=True)
    got = enc(img).cpu().numpy()
    assert got.shape == (128, 64)
    assert np.abs(got - want).max() <= 5e-4
    safe = np.abs(want) > 2e-3
    assert np.array_equal(got[safe] > 0, want[safe] > 0)
    # batch invariance: the same image gives the same code whatever batch it travels in
    again = enc(img[40:52]).cpu().numpy()
    assert np.abs(again - got[40:52]).max() <= 1e-6


@pytest.mark.parametrize("wh,conv", [(64, "tf32x3"), (64, "fp32"), (16, "tf32x3")])
def test_encoder_other_image_sizes(wh, conv):
    """DATA.WIDTH_HEIGHT: 64 (config/nuswide_step_1.yaml:12) -- legacy bilinear 64 -> 256 (scale 1/4) instead of 32 -> 256."""
    from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights
    from oracle import alexnet_oracle

    n = 5
    yy, xx = np.meshgrid(np.arange(wh), np.arange(wh), indexing="ij")
    img = np.stack([np.stack([(7 * yy + 3 * xx + 11 * i) % 256, (5 * yy * (i + 1) + xx) % 256, (yy * xx + 40 * i) % 256]) for i in range(n)]).astype(np.uint8)
    img[3:] = np.random.default_rng(2).integers(0, 256, img[3:].shape, dtype=np.uint8)
    w = AlexNetWeights.synthetic(48, seed=13)
    want = alexnet_oracle.encode(img.reshape(n, -1), w.tensors, wh, lrn=True)
    got = AlexNetHashEncoder(w, lrn=True, conv=conv)(img.reshape(n, -1)).cpu().numpy()
    # "fp32" = CUDA-core convolutions + plain-TF32 dense layers (10-bit mantissa): the looser bound of DESIGN.md section 2
    assert np.abs(got - want).max() <= (5e-4 if conv == "tf32x3" else 5e-3)


@pytest.mark.parametrize("wh,lrn,conv", [(32, True, "tf32x3"), (32, False, "fp32"), (64, True, "tf32x3")])
def test_fused_first_stage_equals_the_separate_kernels(wh, lrn, conv):
    """csrc/encoder_stage1.cu composes resize + crop + flip + conv1 into effective filters (exact algebra) and fuses pool1 + LRN:
    against the separate kernels the outputs may differ only by fp32 summation order (and the separate tensor-core conv1's
    own rounding), in the deterministic and in the stochastic mode; structured images expose offset / flip / phase mistakes."""
    from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights

    n = 7
    yy, xx = np.meshgrid(np.arange(wh), np.arange(wh), indexing="ij")
    img = np.stack([np.stack([(7 * yy + 3 * xx + 11 * i) % 256, (5 * yy * (i + 1) + xx) % 256, (yy * xx + 40 * i) % 256]) for i in range(n)]).astype(np.uint8)
    img[4:] = np.random.default_rng(3).integers(0, 256, img[4:].shape, dtype=np.uint8)
    w = AlexNetWeights.synthetic(64, seed=17)
    for deterministic in (True, False):
        a = AlexNetHashEncoder(w, lrn=lrn, conv=conv, fused_stage1=True, deterministic=deterministic, seed=5)(img.reshape(n, -1)).cpu().numpy()
        b = AlexNetHashEncoder(w, lrn=lrn, conv=conv, fused_stage1=False, deterministic=deterministic, seed=5)(img.reshape(n, -1)).cpu().numpy()
        # the plain-TF32 dense layers of conv="fp32" amplify the fp32-rounding difference of conv1 (10-bit mantissa operands)
        assert np.abs(a - b).max() <= (2e-4 if conv == "tf32x3" else 2e-3), (deterministic, np.abs(a - b).max())


def test_encoder_stages_match_oracle():
    """Stage-by-stage check of the fused prep kernel (normalize + legacy bilinear + 10-crop + mean) through conv1."""
    import torch
    from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights
    from oracle import alexnet_oracle as ao

    # a structured image exposes flip / crop-offset / interpolation mistakes that noise would hide
    n, wh = 3, 32
    yy, xx = np.meshgrid(np.arange(wh), np.arange(wh), indexing="ij")
    img = np.stack([np.stack([(7 * yy + 3 * xx + 11 * i) % 256, (5 * yy * (i + 1) + xx) % 256, (yy * xx + 40 * i) % 256]) for i in range(n)]).astype(np.uint8)
    w = AlexNetWeights.synthetic(32, seed=9)
    want = ao.encode(img.reshape(n, -1), w.tensors, wh, lrn=True)
    got = AlexNetHashEncoder(w, lrn=True)(img.reshape(n, -1)).cpu().numpy()
    assert np.abs(got - want).max() <= 5e-3


def test_evaluate_loop_and_cli_surface(tmp_path):
    """main.py:151-164 mirror on a synthetic dataloader: forward_all truncates the wrapped last batch, evaluate returns the
    same mAP as the oracle metric on the encoder's outputs."""
    import torch
    from types import SimpleNamespace as NS
    from hashgan_b200.config import get_default_config
    from hashgan_b200.dataloader import SyntheticDataloader
    from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights
    from hashgan_b200.evaluate import evaluate, forward_all
    from oracle import maps_oracle

    cfg = get_default_config()
    cfg.merge_from_list(["MODEL.HASH_DIM", 32, "DATA.DB_SIZE", 300, "DATA.TEST_SIZE", 40, "DATA.MAP_R", 300, "TRAIN.BATCH_SIZE", 64])
    enc = AlexNetHashEncoder(AlexNetWeights.synthetic(32, seed=1), lrn=False)
    dl = SyntheticDataloader(64, 32, 10, {"database": 300, "test": 40}, seed=4)
    np.random.seed(0)
    db = forward_all(enc, dl.db_gen, 300, cfg)
    assert tuple(db.output.shape) == (300, 32) and db.label.shape == (300, 10)
    val = evaluate(enc, dl, cfg)          # deterministic mode: evaluate() seeds the loader's shuffle with EVAL.SEED
    # oracle metric on the sign codes of the same encoder outputs (same permutation: same seed)
    np.random.seed(cfg.EVAL.SEED)
    db2 = forward_all(enc, dl.db_gen, 300, cfg)
    q2 = forward_all(enc, dl.test_gen, 40, cfg)
    codes = lambda t: np.where(t.cpu().numpy() > 0, 1.0, -1.0).astype(np.float32)
    ref = maps_oracle.OracleMAPs(300, tie="stable").get_maps_by_feature(NS(output=codes(db2.output), label=db2.label),
                                                                          NS(output=codes(q2.output), label=q2.label))
    assert abs(val - ref) <= 1e-12
    # EVAL.PRECISION_RECALL: the same ranking also yields precision@R / recall@R (R == DB_SIZE here: everything is retrieved)
    pr = evaluate(enc, dl, cfg, precision_recall=True)
    assert pr["mAP"] == val and pr["R"] == 300 and abs(pr["recall"] - 1.0) <= 1e-12
    assert abs(pr["precision"] - np.mean(pr["per_query"]["total"] / 300.0)) <= 1e-12


def test_tensor_core_convolution_option():
    """conv="tf32": conv1-5 as implicit GEMM on tcgen05 with plain TF32 operands.  Same graph, TF32 rounding in every layer:
    looser bound."""
    from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights
    from oracle import alexnet_oracle

    w = AlexNetWeights.synthetic(64, seed=3)
    img = np.random.default_rng(5).integers(0, 256, (9, 3 * 32 * 32), dtype=np.uint8)
    want = alexnet_oracle.encode(img, w.tensors, 32, lrn=True)
    got = AlexNetHashEncoder(w, lrn=True, conv_tf32=True)(img).cpu().numpy()
    ref = AlexNetHashEncoder(w, lrn=True, conv="fp32")(img).cpu().numpy()
    assert np.abs(got - want).max() <= 3e-2
    assert np.abs(got - ref).max() <= 3e-2
    margin = np.abs(want) > 0.1
    assert np.array_equal(got[margin] > 0, want[margin] > 0)


def test_stochastic_mode_matches_the_oracle_with_the_same_draws():
    """EVAL.DETERMINISTIC False = the reference's stochastic eval graph (noise main.py:147, dropout at eval
    architecture.py:369,377).  The generator is counter based, so the same draws are handed to the fp32 oracle."""
    from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights, stochastic_draws
    from oracle import alexnet_oracle

    w = AlexNetWeights.synthetic(48, seed=4)
    img = np.random.default_rng(6).integers(0, 256, (5, 3 * 32 * 32), dtype=np.uint8)
    enc = AlexNetHashEncoder(w, lrn=True, deterministic=False, seed=123)
    got1 = enc(img).cpu().numpy()
    s1 = enc.last_seed
    got2 = enc(img).cpu().numpy()
    s2 = enc.last_seed
    assert s1 != s2 and np.abs(got1 - got2).max() > 1e-3            # a fresh draw per call
    det = AlexNetHashEncoder(w, lrn=True)(img).cpu().numpy()
    assert np.abs(got1 - det).max() > 1e-3                            # and not the deterministic graph
    for seed, got in ((s1, got1), (s2, got2)):
        noise, keep6, keep7 = stochastic_draws(seed, len(img), 32)
        assert 0.0 <= noise.min() and noise.max() < 1 / 128 and abs(noise.mean() - 1 / 256) < 2e-4
        assert abs(keep6.mean() - 0.5) < 0.01 and abs(keep7.mean() - 0.5) < 0.01 and (keep6 != keep7).mean() > 0.4
        want = alexnet_oracle.encode(img, w.tensors, 32, lrn=True, noise=noise, keep6=keep6, keep7=keep7)
        assert np.abs(got - want).max() <= 5e-3
    again = AlexNetHashEncoder(w, lrn=True, deterministic=False, seed=123)(img).cpu().numpy()
    assert np.array_equal(again, got1)                                # same seed, same call index: same output


def test_error_compensated_tensor_core_path_is_fp32_grade():
    """conv='tf32x3' (the default): conv1-5 AND fc6-8 as implicit GEMM on tcgen05 with hi/lo-split TF32 operands (3 MMAs
    per product).  The whole encoder then agrees with the fp32 oracle an order of magnitude better than the paths whose
    dense layers use plain TF32."""
    from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights
    from oracle import alexnet_oracle

    w = AlexNetWeights.synthetic(64, seed=3)
    img = np.random.default_rng(5).integers(0, 256, (9, 3 * 32 * 32), dtype=np.uint8)
    want = alexnet_oracle.encode(img, w.tensors, 32, lrn=True)
    err = {c: np.abs(AlexNetHashEncoder(w, lrn=True, conv=c)(img).cpu().numpy() - want).max() for c in ("tf32x3", "fp32", "tf32")}
    assert err["tf32x3"] <= 5e-4          # fp32 summation-order noise only
    assert err["fp32"] <= 5e-3            # fp32 convolutions + plain-TF32 dense layers (the north_star's split)
    assert err["tf32"] <= 3e-2            # plain TF32 everywhere
    assert err["tf32x3"] < 0.25 * err["fp32"]
    got = AlexNetHashEncoder(w, lrn=True)(img).cpu().numpy()      # default == tf32x3
    margin = np.abs(want) > 2e-3
    assert np.array_equal(got[margin] > 0, want[margin] > 0)      # code bits equal except where |h| < 2e-3
