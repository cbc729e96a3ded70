"""Import-compatible mirror of the reference's `creste` package for the hot path.

Put `creste_public_b200/` on sys.path *before* the reference tree and the reference's train
scripts (`creste/train_ssc.py`, `creste/train_traversability.py`, `scripts/runtime/compile.py`)
resolve `creste.models.*` / `creste.utils.loss_utils` to these sm_100a-backed modules instead of
the PyTorch-eager ones.  Same class names, constructor / forward signatures, output-dict keys
and state_dict layout as the reference.
"""
