"""Config surface of the reference, restated on PyYAML (yacs 0.1.4 is not installable offline).

Reference: lib/config.py:4-52 (defaults tree MODEL / DATA / TRAIN), lib/config.py:55-68
(update_and_inference_config: merge yaml -> derive IMAGE_DIR / MODEL_DIR / LOG_DIR / OUTPUT_DIM ->
makedirs -> freeze -> return the same global object), config/*.yaml.  The yacs behaviour that the
reference relies on (pinned yacs==0.1.4, environment.yml:97) is reproduced: attribute access, unknown
key -> KeyError, string values decoded with ast.literal_eval ("1e-4" -> float), replacement type must
equal the default's type (list<->tuple coercion only) else ValueError, frozen nodes reject writes.
"parity unpinned": yacs is absent from /root/reference and from this image, so these semantics are
restated from its published behaviour and pinned only by tests/test_config.py.

Additive keys (defaults only, so the reference's yamls still load unchanged): the EVAL node.
"""
from __future__ import annotations

import ast
import copy
import os

import yaml

__all__ = ["CfgNode", "config", "get_default_config", "update_and_inference_config"]


class CfgNode(dict):
    IMMUTABLE = "__immutable__"

    def __init__(self, init_dict=None):
        super().__init__()
        self.__dict__[CfgNode.IMMUTABLE] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # attribute-style access, used everywhere in the reference (cfg.DATA.MAP_R, main.py:164)
    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        if name in self.__dict__:
            raise AttributeError(f"Invalid attempt to modify internal CfgNode state: {name}")
        self[name] = value

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__[CfgNode.IMMUTABLE] = self.__dict__[CfgNode.IMMUTABLE]
        return out

    def is_frozen(self):
        return self.__dict__[CfgNode.IMMUTABLE]

    def freeze(self):
        self._immutable(True)

    def defrost(self):
        self._immutable(False)

    def _immutable(self, flag):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._immutable(flag)

    def clone(self):
        return copy.deepcopy(self)

    def dump(self, **kwargs):
        def plain(node):
            return {k: plain(v) if isinstance(v, CfgNode) else v for k, v in node.items()}

        return yaml.safe_dump(plain(self), **kwargs)

    # -- merging (yacs: merge_from_file -> _merge_a_into_b) -------------------------------------
    def merge_from_file(self, cfg_filename):
        with open(cfg_filename, "r") as fh:
            loaded = yaml.safe_load(fh) or {}
        self.merge_from_other_cfg(CfgNode(loaded))

    def merge_from_other_cfg(self, other):
        if self.is_frozen():
            raise AttributeError("Attempted to merge into an immutable CfgNode")
        _merge_a_into_b(other, self, [])

    def merge_from_list(self, cfg_list):
        if len(cfg_list) % 2 != 0:
            raise AssertionError(f"Override list has odd length: {cfg_list}; it must be a list of pairs")
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            parts = full_key.split(".")
            for p in parts[:-1]:
                if p not in node:
                    raise KeyError(f"Non-existent key: {full_key}")
                node = node[p]
            if parts[-1] not in node:
                raise KeyError(f"Non-existent key: {full_key}")
            value = _decode(v)
            node[parts[-1]] = _check_and_coerce(value, node[parts[-1]], parts[-1], full_key)


def _decode(v):
    """yacs _decode_cfg_value: dicts become nodes, strings go through literal_eval, the rest is kept."""
    if isinstance(v, dict):
        return CfgNode(v)
    if not isinstance(v, str):
        return v
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError):
        return v


def _check_and_coerce(replacement, original, key, full_key):
    ot, rt = type(original), type(replacement)
    if rt == ot or original is None:
        return replacement
    for from_t, to_t in ((list, tuple), (tuple, list)):
        if rt == from_t and ot == to_t:
            return to_t(replacement)
    raise ValueError(f"Type mismatch ({ot} vs. {rt}) with values ({original} vs. {replacement}) for config key: {full_key}")


def _merge_a_into_b(a, b, key_list):
    for k, v_ in a.items():
        full_key = ".".join(key_list + [k])
        if k not in b:
            raise KeyError(f"Non-existent config key: {full_key}")
        v = _decode(copy.deepcopy(v_))
        v = _check_and_coerce(v, b[k], k, full_key)
        if isinstance(v, CfgNode):
            _merge_a_into_b(v, b[k], key_list + [k])
        else:
            b[k] = v


# Defaults tree: same keys and values as lib/config.py:6-52 (the yaml merge rejects unknown keys, so the
# key set is part of the config surface), plus the additive EVAL node.
_OUT = "./output/cifar10_step_1"
_DEFAULTS = {
    "MODEL": {
        "DIM_G": 128, "DIM_D": 128, "DIM": 64,
        "HASH_DIM": 64,                                  # lib/config.py:10; no shipped yaml overrides it
        "G_ARCHITECTURE": "NORM", "D_ARCHITECTURE": "NORM",   # D: GOOD | NORM | ALEXNET
        "G_PRETRAINED_MODEL_PATH": "", "D_PRETRAINED_MODEL_PATH": "",
        "ALEXNET_PRETRAINED_MODEL_PATH": "./pretrained_models/reference_pretrain.npy",
    },
    "DATA": {
        "USE_DATASET": "cifar10",
        "LIST_ROOT": "./data/cifar10", "DATA_ROOT": "./data_list/cifar10",   # (sic) swapped in the reference, lib/config.py:20-21
        "LABEL_DIM": 10, "DB_SIZE": 54000, "TEST_SIZE": 1000, "WIDTH_HEIGHT": 32, "OUTPUT_DIM": 3 * 32 ** 2,
        "MAP_R": 54000,
        "OUTPUT_DIR": _OUT,
        "IMAGE_DIR": os.path.join(_OUT, "images"), "MODEL_DIR": os.path.join(_OUT, "models"), "LOG_DIR": os.path.join(_OUT, "logs"),
    },
    "TRAIN": {
        "EVALUATE_MODE": False, "BATCH_SIZE": 64, "ITERS": 100000, "CROSS_ENTROPY_ALPHA": 5, "LR": 1e-4, "G_LR": 1e-4,
        "DECAY": True, "N_CRITIC": 5, "EVAL_FREQUENCY": 20000, "CHECKPOINT_FREQUENCY": 2000, "SAMPLE_FREQUENCY": 1000,
        "ACGAN_SCALE": 1.0, "ACGAN_SCALE_FAKE": 1.0,
        "WGAN_SCALE": 1.0,                               # == 0 switches the AlexNet LRN layers on (lib/architecture.py:268,294)
        "WGAN_SCALE_GP": 10.0, "ACGAN_SCALE_G": 0.1, "WGAN_SCALE_G": 1.0, "NORMED_CROSS_ENTROPY": True, "FAKE_RATIO": 1.0,
    },
    "EVAL": {                                            # additive, build-only
        "BINARIZE": True,        # sign() the hash outputs and rank by Hamming distance (the B200 hot path); False = rank the raw
                                 # outputs by inner product exactly as lib/metric.py:13-14 does (hg_ip_map)
        "CONV": "tf32x3",        # conv1-5: "tf32x3" = implicit GEMM on the tensor cores with error-compensated TF32 (hi/lo split,
                                 # fp32-grade accuracy); "fp32" = CUDA cores; "tf32" = plain TF32 operands (fastest, outputs
                                 # within 3e-2 of the fp32 graph, code bits equal where |h| > 0.1)
        "CONV_TF32": False,      # older spelling of CONV: "tf32"
        "TIE_BREAK": "index",    # (distance asc, database row asc) == np.argsort(kind='stable')
        "NUM_GPUS": 1,           # informational: the GPU count comes from the launcher (torchrun --nproc-per-node N main.py ...);
                                 # evaluate() shards database and queries by rows over the ranks it finds
        "DETERMINISTIC": True,   # no de-quantisation noise (main.py:147), no eval-time dropout (architecture.py:369,377);
                                 # False = the reference's stochastic eval graph, draws seeded by EVAL.SEED
        "PRECISION_RECALL": False,  # also print precision@R / recall@R of the same ranking (MAPs.precision_recall; single process)
        "SYNTHETIC": False,      # seeded synthetic images / weights when the data and checkpoints are absent
        "SEED": 0,
    },
}


def get_default_config() -> CfgNode:
    """Fresh copy of the defaults tree."""
    return CfgNode(copy.deepcopy(_DEFAULTS))


# lib/config.py:4: a module-level global that update_and_inference_config mutates and returns
config = get_default_config()


def update_and_inference_config(cfg_file, cfg: CfgNode | None = None, make_dirs: bool = True, opts=None) -> CfgNode:
    """lib/config.py:55-68.  With cfg=None it updates and returns the module-level `config` like the
    reference does.  `opts` (additive): KEY VALUE pairs merged AFTER the yaml file, i.e. command-line overrides win."""
    c = config if cfg is None else cfg
    c.merge_from_file(cfg_file)
    if opts:
        c.merge_from_list(list(opts))
    c.DATA.IMAGE_DIR = os.path.join(c.DATA.OUTPUT_DIR, "images")
    c.DATA.MODEL_DIR = os.path.join(c.DATA.OUTPUT_DIR, "models")
    c.DATA.LOG_DIR = os.path.join(c.DATA.OUTPUT_DIR, "logs")
    c.DATA.OUTPUT_DIM = 3 * (c.DATA.WIDTH_HEIGHT ** 2)
    if make_dirs:
        os.makedirs(c.DATA.IMAGE_DIR, exist_ok=True)
        os.makedirs(c.DATA.MODEL_DIR, exist_ok=True)
        os.makedirs(c.DATA.LOG_DIR, exist_ok=True)
    c.freeze()
    return c
