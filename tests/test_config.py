"""Config surface (lib/config.py + config/*.yaml) restated on PyYAML; yacs semantics pinned here."""
import os
import textwrap

import pytest

from hashgan_b200.config import CfgNode, get_default_config, update_and_inference_config
from tests import helpers

CIFAR_EVAL_YAML = textwrap.dedent("""
    MODEL:
        G_ARCHITECTURE: "NORM"
        D_ARCHITECTURE: "ALEXNET"
        D_PRETRAINED_MODEL_PATH: "./output/x/models/D_9999.ckpt"
    DATA:
        USE_DATASET: "cifar10"  # comment
        LABEL_DIM: 10
        DB_SIZE: 54000
        TEST_SIZE: 1000
        WIDTH_HEIGHT: 32
        MAP_R: 54000
        LIST_ROOT: "./data_list/cifar10"
        DATA_ROOT: "./data/cifar10"
        OUTPUT_DIR: "{out}"
    TRAIN:
        EVALUATE_MODE: True
        BATCH_SIZE: 128
        LR: 1e-4  # PyYAML reads this as the string '1e-4'; yacs literal_evals it back to a float
        G_LR: 0.0
        WGAN_SCALE: 0.0
""")


def _write(tmp_path, text):
    p = tmp_path / "cfg.yaml"
    p.write_text(text)
    return str(p)


def test_eval_yaml_merges_like_yacs(tmp_path):
    out = str(tmp_path / "out")
    cfg = update_and_inference_config(_write(tmp_path, CIFAR_EVAL_YAML.format(out=out)), get_default_config())
    assert cfg.MODEL.D_ARCHITECTURE == "ALEXNET"
    assert cfg.MODEL.HASH_DIM == 64          # not set by the yaml: the default applies
    assert cfg.TRAIN.EVALUATE_MODE is True and cfg.TRAIN.BATCH_SIZE == 128
    assert isinstance(cfg.TRAIN.LR, float) and cfg.TRAIN.LR == 1e-4
    assert cfg.TRAIN.WGAN_SCALE == 0.0
    assert cfg.DATA.MAP_R == 54000 and cfg.DATA.OUTPUT_DIM == 3 * 32 * 32
    # derived directories are recomputed from OUTPUT_DIR and created (lib/config.py:58-65)
    assert cfg.DATA.MODEL_DIR == os.path.join(out, "models")
    for d in (cfg.DATA.IMAGE_DIR, cfg.DATA.MODEL_DIR, cfg.DATA.LOG_DIR):
        assert os.path.isdir(d)
    # frozen (lib/config.py:67)
    with pytest.raises(AttributeError):
        cfg.DATA.MAP_R = 5
    with pytest.raises(AttributeError):
        cfg.MODEL.HASH_DIM = 32


def test_hash_dim_override_and_width(tmp_path):
    cfg = update_and_inference_config(_write(tmp_path, "MODEL:\n    HASH_DIM: 32\nDATA:\n    WIDTH_HEIGHT: 64\n    OUTPUT_DIR: '%s'\n" % (tmp_path / "o")),
                                      get_default_config())
    assert cfg.MODEL.HASH_DIM == 32 and cfg.DATA.OUTPUT_DIM == 3 * 64 * 64


def test_unknown_key_and_type_mismatch(tmp_path):
    with pytest.raises(KeyError):
        get_default_config().merge_from_file(_write(tmp_path, "MODEL:\n    NOT_A_KEY: 1\n"))
    with pytest.raises(ValueError):  # int default, string override
        get_default_config().merge_from_file(_write(tmp_path, "DATA:\n    MAP_R: 'many'\n"))
    with pytest.raises(ValueError):  # float default, int override: yacs 0.1.4 does not coerce int -> float
        get_default_config().merge_from_file(_write(tmp_path, "TRAIN:\n    LR: 1\n"))
    c = get_default_config()
    c.merge_from_list(["DATA.MAP_R", "5000", "EVAL.NUM_GPUS", 8])
    assert c.DATA.MAP_R == 5000 and c.EVAL.NUM_GPUS == 8
    with pytest.raises(KeyError):
        c.merge_from_list(["DATA.NOPE", 1])


def test_node_behaves_like_a_dict_and_clones():
    c = get_default_config()
    assert isinstance(c, dict) and isinstance(c.DATA, CfgNode)
    d = c.clone()
    d.DATA.MAP_R = 1
    assert c.DATA.MAP_R == 54000
    assert "MAP_R: 54000" in c.dump()
    with pytest.raises(AttributeError):
        c.DATA.NOPE


@pytest.mark.skipif(not helpers.have_reference(), reason="/root/reference not mounted (GPU box)")
def test_every_reference_yaml_loads_and_defaults_match(tmp_path):
    """All four shipped yamls merge cleanly, and the defaults tree equals lib/config.py's (parsed, not imported:
    yacs is not installed)."""
    import ast
    import re

    ref = helpers.REFERENCE_DIR
    for name in sorted(os.listdir(os.path.join(ref, "config"))):
        cfg = get_default_config()
        update_and_inference_config(os.path.join(ref, "config", name), cfg, make_dirs=False)
        assert cfg.is_frozen()
    src = open(os.path.join(ref, "lib", "config.py")).read()
    c = get_default_config()
    checked = 0
    for m in re.finditer(r"^config\.(\w+)\.(\w+) = (.+?)(?:\s+#.*)?$", src, flags=re.M):
        node, key, expr = m.groups()
        try:
            want = ast.literal_eval(expr)
        except Exception:
            continue  # derived values (os.path.join, OUTPUT_DIM expression)
        assert c[node][key] == want and type(c[node][key]) is type(want), (node, key)
        checked += 1
    assert checked >= 35


def test_command_line_overrides_win_over_the_yaml(tmp_path):
    """main.py's trailing KEY VALUE pairs are merged after the yaml file (derived keys follow them)."""
    from hashgan_b200.config import get_default_config, update_and_inference_config

    y = tmp_path / "c.yaml"
    y.write_text("DATA:\n    DB_SIZE: 5400\n    WIDTH_HEIGHT: 32\n    OUTPUT_DIR: '%s'\n" % (tmp_path / "out"))
    c = update_and_inference_config(str(y), get_default_config(), opts=["DATA.DB_SIZE", "54000", "DATA.WIDTH_HEIGHT", "64"])
    assert c.DATA.DB_SIZE == 54000 and c.DATA.WIDTH_HEIGHT == 64 and c.DATA.OUTPUT_DIM == 3 * 64 * 64
    c2 = update_and_inference_config(str(y), get_default_config())
    assert c2.DATA.DB_SIZE == 5400 and c2.DATA.OUTPUT_DIM == 3 * 32 * 32
