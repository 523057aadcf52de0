"""Load the UNMODIFIED reference (`/root/reference/ot_vae_lightning`) in the authoring container.

Only `tests/golden/make_golden.py` uses this (to write the committed fixtures).  Nothing that runs on the GPU box imports
it: `/root/reference` does not exist there.  The stub machinery for the absent third-party packages (pytorch_lightning,
torchmetrics, ...) lives in `baseline/ref_loader.py`, shared with `bench.py --impl reference`.
"""
import os
import sys

REF_ROOT = "/root/reference"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "baseline"))
import ref_loader  # noqa: E402


def load_reference():
    return ref_loader.load_reference(REF_ROOT)
