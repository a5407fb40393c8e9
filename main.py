#!/usr/bin/env python
"""CLI of the evaluation path -- same surface as the reference's main.py:253-270:

    python main.py --cfg config/cifar_evaluation.yaml --gpus 0

prints the merged config, writes it to OUTPUT_DIR/config.txt, evaluates and prints `map_val: <float>`
(main.py:197-199).  Only TRAIN.EVALUATE_MODE: True is in scope (training is not part of this build).
Weights: MODEL.ALEXNET_PRETRAINED_MODEL_PATH (.npy dict, lib/architecture.py:199) when present, then overridden by the
TensorFlow checkpoint MODEL.D_PRETRAINED_MODEL_PATH (main.py:187-195; hashgan_b200/tf_checkpoint.py); with
EVAL.SYNTHETIC: True seeded synthetic weights and images stand in for missing files.
"""
import argparse
import os
import sys
from pprint import pprint


def main(cfg):
    from hashgan_b200.dataloader import Dataloader, SyntheticDataloader
    from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights
    from hashgan_b200.evaluate import evaluate

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:  # torchrun: one process per GPU, database and queries sharded by rows (EVAL.NUM_GPUS documents the intent)
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    if not cfg.TRAIN.EVALUATE_MODE:
        raise SystemExit("hashgan_b200 implements the evaluation path only: set TRAIN.EVALUATE_MODE: True")
    if cfg.MODEL.D_ARCHITECTURE != "ALEXNET":
        raise SystemExit("only MODEL.D_ARCHITECTURE: ALEXNET is in scope (cifar_evaluation.yaml:3)")
    npy = cfg.MODEL.ALEXNET_PRETRAINED_MODEL_PATH
    if os.path.exists(npy):
        weights = AlexNetWeights.from_alexnet_npy(npy, cfg.MODEL.HASH_DIM, cfg.EVAL.SEED)
        print("AlexNet weights loaded: {}".format(npy))
    elif cfg.EVAL.SYNTHETIC:
        weights = AlexNetWeights.synthetic(cfg.MODEL.HASH_DIM, cfg.EVAL.SEED)
        print("synthetic AlexNet weights (seed {})".format(cfg.EVAL.SEED))
    else:
        raise SystemExit("{} not found (set EVAL.SYNTHETIC: True for seeded synthetic weights)".format(npy))
    ckpt = cfg.MODEL.D_PRETRAINED_MODEL_PATH
    if len(ckpt) > 0 and os.path.exists(ckpt + ".index"):
        # main.py:193-194: Saver.restore overrides the initial values with the trained discriminator
        names = weights.override_from_tf_checkpoint(ckpt)
        print("discriminator checkpoint restored: {} ({} tensors)".format(ckpt, len(names)))
    elif len(ckpt) > 0 and not cfg.EVAL.SYNTHETIC:
        raise SystemExit("{}.index not found (set EVAL.SYNTHETIC: True to evaluate without the trained checkpoint)".format(ckpt))
    elif len(ckpt) > 0 and rank == 0:
        print("WARNING: MODEL.D_PRETRAINED_MODEL_PATH = {} was NOT found; EVAL.SYNTHETIC is set, so the hash head is evaluated with "
              "its INITIAL weights (map_val is not the trained model's)".format(ckpt), file=sys.stderr)
    encoder = AlexNetHashEncoder(weights, lrn=(cfg.TRAIN.WGAN_SCALE == 0), conv=("tf32" if cfg.EVAL.CONV_TF32 else cfg.EVAL.CONV),
                                 deterministic=bool(cfg.EVAL.DETERMINISTIC), seed=cfg.EVAL.SEED)
    if os.path.isdir(cfg.DATA.DATA_ROOT) and os.path.isdir(cfg.DATA.LIST_ROOT):
        dataloader = Dataloader(cfg.TRAIN.BATCH_SIZE, cfg.DATA.WIDTH_HEIGHT, cfg.DATA.LIST_ROOT, cfg.DATA.DATA_ROOT)
    elif cfg.EVAL.SYNTHETIC:
        dataloader = SyntheticDataloader(cfg.TRAIN.BATCH_SIZE, cfg.DATA.WIDTH_HEIGHT, cfg.DATA.LABEL_DIM,
                                         {"database": cfg.DATA.DB_SIZE, "test": cfg.DATA.TEST_SIZE}, cfg.EVAL.SEED)
    else:
        raise SystemExit("{} / {} not found (set EVAL.SYNTHETIC: True for seeded synthetic images)".format(cfg.DATA.DATA_ROOT, cfg.DATA.LIST_ROOT))
    pr = None
    if cfg.EVAL.PRECISION_RECALL and world == 1:
        pr = evaluate(encoder, dataloader, cfg, precision_recall=True)
        map_val = pr["mAP"]
    else:
        map_val = evaluate(encoder, dataloader, cfg)
    if rank == 0:
        # the reference ranks the raw crop-averaged tanh outputs by inner product (lib/metric.py:13-14); the B200 hot path
        # (EVAL.BINARIZE True, the default) ranks their signs by Hamming distance -- identical on +-1 codes only
        print('ranking: {}'.format('Hamming distance on sign-binarised codes (EVAL.BINARIZE True; set it False for the reference\'s '
                                   'real-valued inner-product ranking)' if cfg.EVAL.BINARIZE else
                                   'real-valued inner product of the raw outputs (EVAL.BINARIZE False, lib/metric.py:13-14)'))
        print('map_val: {}'.format(map_val))
        if pr is not None:
            print('precision@{}: {}'.format(pr["R"], pr["precision"]))
            print('recall@{}: {}'.format(pr["R"], pr["recall"]))
        elif cfg.EVAL.PRECISION_RECALL:
            print('precision/recall: skipped (EVAL.PRECISION_RECALL is computed by the single-process metric; run without torchrun)')
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.path.append(os.getcwd())
    parser = argparse.ArgumentParser(description='HashGAN evaluation on B200')
    parser.add_argument('--cfg', '--config', required=True, type=str, metavar="FILE", help="path to yaml config")
    parser.add_argument('--gpus', default=None, type=str)
    parser.add_argument('opts', nargs=argparse.REMAINDER, help="KEY VALUE overrides, e.g. EVAL.SYNTHETIC True DATA.DB_SIZE 2048")
    args = parser.parse_args()
    under_torchrun = int(os.environ.get("WORLD_SIZE", "1")) > 1
    if args.gpus is not None:
        os.environ["CUDA_VISIBLE_DEVICES"] = args.gpus       # main.py:263
    elif not under_torchrun:
        os.environ["CUDA_VISIBLE_DEVICES"] = "0"              # the reference's default (--gpus 0)
    # under torchrun without --gpus every rank keeps all GPUs visible and picks LOCAL_RANK

    from hashgan_b200.config import config, update_and_inference_config

    config = update_and_inference_config(args.cfg, opts=args.opts)  # KEY VALUE overrides are merged after the yaml file
    if int(os.environ.get("RANK", "0")) == 0:                 # one copy of the config, not one per rank
        pprint(config)
        with open(os.path.join(config.DATA.OUTPUT_DIR, 'config.txt'), 'w') as fh:
            pprint(config, fh)
    sys.exit(main(config))
