"""Host input pipeline: hashgan_b200.dataloader.Dataloader against golden batches produced by the
UNMODIFIED reference loader (oracle/gen_golden_dataloader.py ran /root/reference/lib/dataloader.py): list-file parsing, cv2
decode + INTER_AREA resize, shuffled epochs with the wrap-around last batch, NHWC->NCHW, BGR->RGB, flatten, label rows."""
import os

import numpy as np
import pytest

from tests import helpers

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def folder(tmp_path_factory):
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "dataloader_golden.npz"), allow_pickle=False)
    root = tmp_path_factory.mktemp("imgs")
    os.makedirs(root / "img" / "sub")
    os.makedirs(root / "lists")
    for split in ("database", "test"):
        lines = [str(s) for s in g[f"{split}/lines"]]
        for i, line in enumerate(lines):
            assert cv2.imwrite(str(root / "img" / line.split()[0]), g[f"{split}/img{i}"])   # PNG: lossless
        (root / "lists" / f"{split}.txt").write_text("\n".join(lines) + "\n")
    return g, root


@pytest.mark.parametrize("wh,batch", [(8, 4), (16, 3)])
def test_batches_equal_the_reference_loader(folder, wh, batch):
    from hashgan_b200.dataloader import Dataloader

    g, root = folder
    dl = Dataloader(batch, wh, str(root / "lists"), str(root / "img"))
    np.random.seed(7)  # the seed the golden run used; permutations are drawn in the same order
    for split, gen in (("database", dl.db_gen), ("test", dl.test_gen)):
        n = int(g[f"{split}/n"])
        for epoch in range(2):
            batches = list(gen())
            assert len(batches) == -(-n // batch)
            for k, (data, label) in enumerate(batches):
                want_d, want_l = g[f"out/{wh}_{batch}/{split}/e{epoch}/b{k}/data"], g[f"out/{wh}_{batch}/{split}/e{epoch}/b{k}/label"]
                assert data.shape == (batch, 3 * wh * wh) and data.dtype == np.uint8
                assert np.array_equal(data, want_d), (split, epoch, k)
                assert np.array_equal(np.asarray(label), want_l)


def test_layout_is_rgb_planes(folder):
    """[B, 3*wh*wh] = RGB planes of the INTER_AREA-resized image (lib/dataloader.py:110-113)."""
    from hashgan_b200.dataloader import Dataloader

    g, root = folder
    wh = 8
    dl = Dataloader(5, wh, str(root / "lists"), str(root / "img"))
    np.random.seed(1)
    perm = np.arange(5)
    np.random.shuffle(perm)
    np.random.seed(1)
    data, label = next(iter(dl.test_gen()))
    for row, i in enumerate(perm):
        bgr = cv2.resize(g[f"test/img{i}"], (wh, wh), interpolation=cv2.INTER_AREA)
        assert np.array_equal(data[row].reshape(3, wh, wh), bgr[:, :, ::-1].transpose(2, 0, 1))
        assert label[row].tolist() == [int(v) for v in str(g["test/lines"][i]).split()[1:]]


def test_missing_image_fails_loudly(folder, tmp_path):
    """The reference swallows the error and later breaks in a reshape (lib/dataloader.py:55-69,113); the mirror raises."""
    from hashgan_b200.dataloader import Dataloader

    (tmp_path / "lists").mkdir()
    (tmp_path / "lists" / "database.txt").write_text("nope.png 1 0\n")
    with pytest.raises(FileNotFoundError):
        next(iter(Dataloader(1, 8, str(tmp_path / "lists"), str(tmp_path)).db_gen()))
