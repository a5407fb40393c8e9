"""hashgan_b200 -- B200-native retrieval-evaluation hot path of thuml/HashGAN.

Public surface mirrors the reference for this path only:
    MAPs(R).get_maps_by_feature(database, query)      lib/metric.py:4-24, main.py:164
    config / update_and_inference_config(path)        lib/config.py:4-68
    forward_all / evaluate                            main.py:151-164
Everything numeric runs in libhashgan_b200.so (hand-written sm_100a CUDA behind include/hashgan_b200.h).
"""
from .metric import MAPs, MAPs_CQ, hamming_map_device, pack_rows  # noqa: F401

__version__ = "0.1.0"
