#!/usr/bin/env python3
"""Stand-in for the reference's scripts/inference.py (same import block, CLI flags and calls; see ../README.md)."""
import torch  # noqa: F401  (the reference imports torch first)

import argparse
import logging
import sys
from pathlib import Path

# Add parent directory to path for imports
sys.path.insert(0, str(Path(__file__).parent.parent))

from config import Config
from models import BathymetricPipeline


def parse_args():
    p = argparse.ArgumentParser()
    for flag in ("--input", "--output", "--model"):
        p.add_argument(flag, type=Path, required=True)
    p.add_argument("--config", type=Path, default=None)
    p.add_argument("--tile-size", type=int, default=None)
    p.add_argument("--auto-correct-threshold", type=float, default=None)
    p.add_argument("--review-threshold", type=float, default=None)
    p.add_argument("--min-valid-ratio", type=float, default=None)
    p.add_argument("--export-extras", action="store_true", default=True)
    p.add_argument("--no-export-extras", action="store_false", dest="export_extras")
    p.add_argument("--device", type=str, default="cuda")
    p.add_argument("--vr-bag-mode", type=str, default="resampled")
    p.add_argument("--log-level", type=str, default="INFO")
    return p.parse_args()


def main():
    args = parse_args()
    logging.basicConfig(level=getattr(logging, args.log_level))
    if not args.input.exists() or not args.model.exists():
        sys.exit(1)
    args.output.parent.mkdir(parents=True, exist_ok=True)
    sidecar_cfg = args.model.parent / "config.yaml"
    if args.config and args.config.exists():
        config = Config.load(args.config)
    elif sidecar_cfg.exists():
        config = Config.load(sidecar_cfg)
    else:
        config = Config()
    config.device = args.device
    if args.tile_size:
        config.tile.tile_size = args.tile_size
    if args.min_valid_ratio is not None:
        config.tile.min_valid_ratio = args.min_valid_ratio
    if args.auto_correct_threshold:
        config.inference.auto_correct_threshold = args.auto_correct_threshold
    if args.review_threshold:
        config.inference.review_threshold = args.review_threshold
    pipeline = BathymetricPipeline(config, vr_bag_mode=args.vr_bag_mode)
    pipeline.load_model(args.model)
    results = pipeline.process(input_path=args.input, output_path=args.output, export_extras=args.export_extras)
    return results


if __name__ == "__main__":
    RESULTS = main()
