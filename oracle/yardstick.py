"""TEST INFRASTRUCTURE (build container only) — precision yardstick: how far is the REFERENCE'S OWN bf16-autocast
path (its training regime, tasks/forecasting.py:22) from its fp32 path on the golden cases?  The kernel stack
computes in the same regime (bf16 operands, fp32 accumulation), so this is the natural scale for its parity numbers.
    python -m oracle.yardstick
"""
import sys
import tempfile
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from _fixtures import CASES, config_for, load_case, materialize_llm_dir  # noqa: E402
from oracle import ref_harness as H  # noqa: E402


def main():
    for name in CASES[:6]:
        fix = load_case(name)
        tmp = Path(tempfile.mkdtemp(prefix="yard_"))
        cfg = config_for(fix, materialize_llm_dir(fix, tmp / "llm"))
        ds = H.SyntheticDataset(fix["dataset"]["n_features"], fix["dataset"]["n_classes"], fix["dataset"]["description"])
        model = H.build_reference_model(cfg, ds)
        model.load_state_dict(fix["adapters"], strict=False)
        model.eval()
        with torch.no_grad():
            ref = model(dict(fix["inputs"]))
            with torch.autocast("cpu", dtype=torch.bfloat16):
                amp = model(dict(fix["inputs"])).float()
        err = ((amp - ref).norm() / ref.norm()).item()
        print(f"{name}: reference bf16-autocast vs reference fp32, relative L2 of the output = {err:.2e}")


if __name__ == "__main__":
    main()
