"""Golden fixture for Trainer.validate (reference trainer/Trainer.py:247-308): runs the UNMODIFIED reference on CPU.

    python tests/golden/make_golden_validate.py      # authoring container only

Recipe: the Trainer of make_golden.py (seed 0, G before D), one train step at 2+1+1 patches of 32^3 drawn from
torch.Generator().manual_seed(21), then validate() with val_iterations = 2 on (64, 64, 32) patches, batch 2 per scan
type, drawn from the same generator.  The logged validation losses are captured from logger.log_loss.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import cgan_oracle as O  # noqa: E402
import make_golden as MG  # noqa: E402


def main():
    ref = ref_shim.load()
    T = ref_shim.load_trainer()
    torch.set_num_threads(8)
    tr, _ = MG.make_trainer(ref, T, (32, 32, 32), 2)
    tr.val_iterations = 2
    tr.val_every = 400
    gen = torch.Generator().manual_seed(21)
    patch = (32, 32, 32)
    opt = O.synthetic_patches(gen, (2, 1, *patch))
    low = O.synthetic_patches(gen, (1, 1, *patch))
    high = O.synthetic_patches(gen, (1, 1, *patch))
    ml = O.synthetic_masks(gen, (1, 1, *patch))
    mh = O.synthetic_masks(gen, (1, 1, *patch))
    tr.train_step([dict(data=opt, seg=torch.zeros_like(opt, dtype=torch.bool), name=["o"] * 2),
                   dict(data=low, seg=ml, name=["l"]), dict(data=high, seg=mh, name=["h"])], 0)
    vpatch = (64, 64, 32)
    vb = [tuple(O.synthetic_patches(gen, (2, 1, *vpatch)) for _ in range(3)) for _ in range(2)]
    captured = {}

    class L:
        @staticmethod
        def log_loss(losses, iteration, stage):
            captured.update({k: float(v) for k, v in losses.items()})

    tr.logger_interface.logger = L()
    ScanType = ref.alias.ScanType
    order = [s.value for s in ScanType]
    loaders = {order[0]: iter([dict(data=b[0], name=["o"] * 2) for b in vb]), order[1]: iter([dict(data=b[1], name=["l"] * 2) for b in vb]),
               order[2]: iter([dict(data=b[2], name=["h"] * 2) for b in vb])}
    tr.validate(loaders, 400)
    assert tr.generator.training and tr.critic.training
    np.savez_compressed(HERE / "validate_32.npz", val=np.array([captured[k] for k in ("D", "G", "sim")], dtype=np.float64))
    print("validate golden [D, G, sim]:", captured)


if __name__ == "__main__":
    main()
