"""Host-side pieces of the encoder that need no GPU: the counter-based generator of the stochastic mode (it must stay
bit-identical to hg_mix in csrc/encoder.cu -- the GPU test feeds its draws to the oracle) and the weight contract."""
import numpy as np
import pytest

from hashgan_b200.encoder import CONV_SHAPES, WEIGHT_NAMES, AlexNetWeights, _mix, stochastic_draws


def test_generator_known_answers():
    # seed 0 is the published splitmix64 sequence: 0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F
    assert [int(x) for x in _mix(0, np.arange(3))] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    assert [int(x) for x in _mix(12345, np.array([0, 1, 1 << 40]))] == [0x22118258A9D111A0, 0x346EDCE5F713F8ED, 0x62EC45759FCCD821]
    noise, keep6, keep7 = stochastic_draws(7, 1, 4)
    assert np.allclose(noise[0, :2], [0.0029440815560519695, 0.006381911225616932], rtol=0, atol=1e-9)
    assert keep6[0, :8].astype(int).tolist() == [0, 0, 0, 0, 1, 1, 1, 1]
    assert keep7[0, :8].astype(int).tolist() == [1, 0, 0, 0, 1, 1, 1, 1]


def test_draw_statistics_match_the_reference_distributions():
    noise, keep6, keep7 = stochastic_draws(99, 8, 32)
    assert noise.shape == (8, 3072) and keep6.shape == keep7.shape == (80, 4096)
    assert noise.min() >= 0.0 and noise.max() < 1 / 128                       # tf.random_uniform(0, 1/128), main.py:147
    assert abs(noise.mean() - 1 / 256) < 1e-4 and abs(noise.var() - (1 / 128) ** 2 / 12) < 1e-7
    assert abs(keep6.mean() - 0.5) < 5e-3 and abs(keep7.mean() - 0.5) < 5e-3   # keep_prob 0.5, architecture.py:369,377
    assert abs((keep6 == keep7).mean() - 0.5) < 5e-3                           # the two layers draw independently
    n2, _, _ = stochastic_draws(100, 8, 32)
    assert abs(np.corrcoef(noise.ravel(), n2.ravel())[0, 1]) < 0.02            # and so do different seeds


def test_weight_contract():
    w = AlexNetWeights.synthetic(48, seed=0)
    assert set(w.tensors) == set(WEIGHT_NAMES(48))
    assert w.tensors["discriminator.conv2.weights"].shape == CONV_SHAPES["conv2"] == (5, 5, 48, 256)
    assert w.tensors["discriminator.ACGANOutput.W"].shape == (4096, 48)
    bad = dict(w.tensors)
    bad["discriminator.fc6.weights"] = np.zeros((4096, 9216), np.float32)
    with pytest.raises(ValueError, match="expected shape"):
        AlexNetWeights(bad, 48)
    del bad["discriminator.fc6.weights"]
    with pytest.raises(KeyError):
        AlexNetWeights(bad, 48)


def test_alexnet_npy_import_follows_the_reference_dict_layout(tmp_path):
    """lib/architecture.py:199: net_data = np.load(path, encoding='latin1').item(); net_data[layer][0|1] = weights | biases
    for conv1-5, fc6, fc7 (fc8 is replaced by the hash layer, lib/ops.py `linear` initialisation)."""
    rng = np.random.default_rng(2)
    net = {}
    for layer, shape in list(CONV_SHAPES.items()) + [("fc6", (9216, 4096)), ("fc7", (4096, 4096))]:
        if layer.startswith("fc"):
            shape = (shape[0] // 64, shape[1] // 64)  # keep the fixture small; shapes are validated below with full-size arrays
        net[layer] = [rng.standard_normal(shape).astype(np.float32), rng.standard_normal(shape[-1]).astype(np.float32)]
    net["fc8"] = [np.zeros((4, 1000), np.float32), np.zeros(1000, np.float32)]  # present in bvlc_alexnet.npy, unused by the hash head
    path = str(tmp_path / "reference_pretrain.npy")
    np.save(path, np.array(net, dtype=object), allow_pickle=True)
    with pytest.raises(ValueError, match="fc6"):          # the reduced fc fixture must be rejected: shapes are part of the contract
        AlexNetWeights.from_alexnet_npy(path, 64)
    net["fc6"] = [np.zeros((9216, 4096), np.float32), np.ones(4096, np.float32)]
    net["fc7"] = [np.zeros((4096, 4096), np.float32), np.ones(4096, np.float32)]
    np.save(path, np.array(net, dtype=object), allow_pickle=True)
    w = AlexNetWeights.from_alexnet_npy(path, 64, seed=3)
    for layer in CONV_SHAPES:
        assert np.array_equal(w.tensors[f"discriminator.{layer}.weights"], net[layer][0])
        assert np.array_equal(w.tensors[f"discriminator.{layer}.biases"], net[layer][1])
    assert np.array_equal(w.tensors["discriminator.fc7.biases"], np.ones(4096, np.float32))
    W8 = w.tensors["discriminator.ACGANOutput.W"]
    lim = np.sqrt(2.0 / (4096 + 64)) * np.sqrt(3.0)                      # Glorot uniform, lib/ops.py:213-218
    assert W8.shape == (4096, 64) and np.abs(W8).max() <= lim and W8.std() > 0.4 * lim
    assert not w.tensors["discriminator.ACGANOutput.b"].any()


def test_periodic_evaluate_follows_the_reference_schedule(tmp_path, capsys):
    """main.py:236-240: evaluate every TRAIN.EVAL_FREQUENCY iterations and on the last one, print `map_val: ...`, log the
    scalar `mAP_feature` at that step.  The metric is injected, so no GPU is involved."""
    import json
    from types import SimpleNamespace as NS

    import torch

    from hashgan_b200 import evaluate as ev
    from hashgan_b200.dataloader import SyntheticDataloader

    cfg = NS(MODEL=NS(HASH_DIM=8), DATA=NS(DB_SIZE=20, TEST_SIZE=6, LABEL_DIM=3, MAP_R=5), TRAIN=NS(BATCH_SIZE=4, EVAL_FREQUENCY=5, ITERS=12),
             EVAL=NS(SEED=1, BINARIZE=True, DETERMINISTIC=True))
    loader = SyntheticDataloader(4, 2, 3, {"database": 20, "test": 6}, seed=2)
    enc = lambda image: torch.from_numpy(np.asarray(image, dtype=np.float32)[:, :8] / 255.0 - 0.5)  # noqa: E731
    calls = []

    class FakeMetric:
        def get_maps_by_feature(self, db, q):
            assert tuple(db.output.shape) == (20, 8) and tuple(q.output.shape) == (6, 8)
            calls.append(1)
            return np.float64(0.25 + 0.125 * len(calls))

    log = ev.ScalarLog(str(tmp_path / "logs"))
    got = [ev.periodic_evaluate(it, enc, loader, cfg, summary_writer=log, metric=FakeMetric()) for it in range(cfg.TRAIN.ITERS)]
    due = [it for it, v in enumerate(got) if v is not None]
    assert due == [4, 9, 11]                                       # (it + 1) % 5 == 0, and the last iteration
    lines = [json.loads(ln) for ln in open(log.path)]
    assert [(ln["tag"], ln["step"]) for ln in lines] == [("mAP_feature", 4), ("mAP_feature", 9), ("mAP_feature", 11)]
    assert [ln["value"] for ln in lines] == [0.375, 0.5, 0.625]
    printed = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("map_val: ")]
    assert printed == ["map_val: 0.375", "map_val: 0.5", "map_val: 0.625"]
    s = ev.scalar_summary("mAP_feature", 0.5)
    assert (s.tag, s.simple_value) == ("mAP_feature", 0.5)
