"""TensorFlow-1.x V2 checkpoint reader (hashgan_b200/tf_checkpoint.py): replaces tf.train.Saver.restore of
main.py:187-195.  TensorFlow is absent here (parity unpinned): the tests pin the published pieces of the format --
CRC-32C known answers, the LevelDB mask, the table magic, varint / protobuf wire encodings -- and check that
corruption anywhere is detected through the checksums the format carries."""
import os
import struct

import numpy as np
import pytest

from hashgan_b200 import tf_checkpoint as tfc


def test_crc32c_known_answers():
    # RFC 3720 B.4 / the LevelDB crc32c_test vectors
    assert tfc.crc32c(b"123456789") == 0xE3069283
    assert tfc.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert tfc.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tfc.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfc.crc32c(b"hello world") == tfc.crc32c(b" world", tfc.crc32c(b"hello"))
    # LevelDB masking: rotate right by 15, add 0xa282ead8; must round-trip and differ from the raw value
    c = tfc.crc32c(b"foo")
    assert tfc._mask(c) != c and tfc._unmask(tfc._mask(c)) == c
    assert tfc._mask(0) == 0xA282EAD8


def test_native_crc_matches_python():
    from hashgan_b200 import _native

    lib = _native.lib()
    rng = np.random.default_rng(0)
    for n in (0, 1, 3, 4, 5, 63, 4097, 100001):
        buf = rng.integers(0, 256, n, dtype=np.uint8)
        want = 0xFFFFFFFF
        for b in buf.tobytes():
            want = tfc._TABLE[(want ^ b) & 0xFF] ^ (want >> 8)
        want ^= 0xFFFFFFFF
        got = lib.hg_crc32c(buf.ctypes.data if n else None, n, 0)
        assert got == want


def _weights(rng):
    return {
        "discriminator.conv1.weights": rng.standard_normal((11, 11, 3, 96)).astype(np.float32),
        "discriminator.conv1.biases": rng.standard_normal((96,)).astype(np.float32),
        "discriminator.ACGANOutput.W": rng.standard_normal((4096, 64)).astype(np.float32),
        "discriminator.ACGANOutput.b": np.zeros((64,), np.float32),
        "discriminator.Output.W": rng.standard_normal((4096, 1)).astype(np.float32),
        "generator.Input.W": rng.standard_normal((7, 5)).astype(np.float64),
        "global_step": np.asarray(20000, dtype=np.int64),
        "beta1_power": np.asarray(0.5, dtype=np.float32),
    }


@pytest.mark.parametrize("per_block", [1, 3, 100])
def test_round_trip_and_layout(tmp_path, per_block):
    rng = np.random.default_rng(1)
    w = _weights(rng)
    prefix = str(tmp_path / "D_20000.ckpt")
    tfc.write_checkpoint(prefix, w, entries_per_block=per_block)
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw) >= 48
    assert os.path.getsize(prefix + ".data-00000-of-00001") == sum(a.nbytes for a in w.values())
    listed = tfc.list_variables(prefix)
    assert set(listed) == set(w)
    assert listed["discriminator.conv1.weights"] == (np.dtype(np.float32), (11, 11, 3, 96))
    assert listed["global_step"] == (np.dtype(np.int64), ())
    got = tfc.read_checkpoint(prefix)
    for k, a in w.items():
        assert got[k].dtype == a.dtype and got[k].shape == a.shape and np.array_equal(got[k], a)
    sub = tfc.read_checkpoint(prefix, ["discriminator.conv1.biases"])
    assert list(sub) == ["discriminator.conv1.biases"]
    with pytest.raises(KeyError):
        tfc.read_checkpoint(prefix, ["discriminator.fc6.weights"])


def test_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(2)
    w = _weights(rng)
    prefix = str(tmp_path / "D.ckpt")
    tfc.write_checkpoint(prefix, w)
    data_path, index_path = prefix + ".data-00000-of-00001", prefix + ".index"
    blob = bytearray(open(data_path, "rb").read())
    blob[1000] ^= 0x01
    open(data_path, "wb").write(bytes(blob))
    with pytest.raises(tfc.CheckpointError, match="tensor checksum"):
        tfc.read_checkpoint(prefix)
    assert tfc.read_checkpoint(prefix, verify=False)  # explicit opt-out still reads
    blob[1000] ^= 0x01
    open(data_path, "wb").write(bytes(blob[:-8]))
    with pytest.raises(tfc.CheckpointError, match="truncated"):
        tfc.read_checkpoint(prefix)
    idx = bytearray(open(index_path, "rb").read())
    idx[10] ^= 0x40
    open(index_path, "wb").write(bytes(idx))
    with pytest.raises(tfc.CheckpointError, match="block checksum"):
        tfc.list_variables(prefix)
    idx[10] ^= 0x40
    idx[-1] ^= 0xFF
    open(index_path, "wb").write(bytes(idx))
    with pytest.raises(tfc.CheckpointError, match="magic"):
        tfc.list_variables(prefix)
    with pytest.raises(FileNotFoundError):
        tfc.read_checkpoint(str(tmp_path / "missing.ckpt"))


def test_restore_order_of_the_reference(tmp_path):
    """main.py:187-195: initial (.npy / synthetic) values first, then the checkpoint overrides what it holds."""
    from hashgan_b200.encoder import AlexNetWeights

    base = AlexNetWeights.synthetic(48, seed=1)
    trained = AlexNetWeights.synthetic(48, seed=2)
    ck = {k: v for k, v in trained.tensors.items() if "conv" in k or "ACGANOutput" in k}
    ck["discriminator.Output.W"] = np.zeros((4096, 1), np.float32)      # the WGAN head: in the checkpoint, not on the eval path
    ck["discriminator.conv1.weights/Adam"] = np.zeros((11, 11, 3, 96), np.float32)
    prefix = str(tmp_path / "D_1.ckpt")
    tfc.write_checkpoint(prefix, ck)
    # fc6 / fc7 are missing: tf.train.Saver.restore (main.py:193) raises NotFoundError, so must we
    with pytest.raises(KeyError, match="lacks 4 variable"):
        AlexNetWeights.synthetic(48, seed=1).override_from_tf_checkpoint(prefix)
    names = base.override_from_tf_checkpoint(prefix, allow_partial=True)
    assert sorted(names) == sorted(k for k in ck if k in trained.tensors)
    for k in base.tensors:
        src = trained if k in names else AlexNetWeights.synthetic(48, seed=1)
        assert np.array_equal(base.tensors[k], src.tensors[k])
    full = dict(trained.tensors)
    full["discriminator.Output.W"] = ck["discriminator.Output.W"]
    tfc.write_checkpoint(prefix, full)
    strict = AlexNetWeights.synthetic(48, seed=1)
    assert sorted(strict.override_from_tf_checkpoint(prefix)) == sorted(trained.tensors)   # complete checkpoint: strict mode passes
    assert all(np.array_equal(strict.tensors[k], trained.tensors[k]) for k in trained.tensors)
    bad = dict(full)
    bad["discriminator.ACGANOutput.W"] = np.zeros((4096, 64), np.float32)  # HASH_DIM mismatch
    tfc.write_checkpoint(prefix, bad)
    with pytest.raises(ValueError, match="checkpoint shape"):
        AlexNetWeights.synthetic(48, seed=1).override_from_tf_checkpoint(prefix)


def test_round_trip_property(tmp_path):
    """Randomised round trips (hypothesis): arbitrary variable names (prefix compression, restart points, multi-block
    index), ranks 0..4, every supported dtype."""
    from hypothesis import HealthCheck, given, settings, strategies as st

    dtypes = [np.float32, np.float64, np.int32, np.int64, np.uint8, np.int8, np.int16, np.float16, np.bool_]
    names = st.text(alphabet="abcdefghij./_0123456789", min_size=1, max_size=24)
    shapes = st.lists(st.integers(0, 5), min_size=0, max_size=4)
    counter = [0]

    @settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.dictionaries(names, st.tuples(st.sampled_from(range(len(dtypes))), shapes, st.integers(0, 2 ** 31)), min_size=1, max_size=20),
           st.integers(1, 7))
    def run(spec, per_block):
        tensors = {}
        for name, (di, shape, seed) in spec.items():
            rng = np.random.default_rng(seed)
            a = (rng.normal(size=shape) * 100).astype(dtypes[di]) if dtypes[di] is not np.bool_ else rng.random(shape) < 0.5
            tensors[name] = np.asarray(a)
        counter[0] += 1
        prefix = str(tmp_path / f"p{counter[0]}.ckpt")
        tfc.write_checkpoint(prefix, tensors, entries_per_block=per_block)
        got = tfc.read_checkpoint(prefix)
        assert set(got) == set(tensors)
        for k, a in tensors.items():
            assert got[k].dtype == a.dtype and got[k].shape == a.shape and np.array_equal(got[k], a)
        assert {k: v[1] for k, v in tfc.list_variables(prefix).items()} == {k: tuple(a.shape) for k, a in tensors.items()}

    run()
