"""One MKGformer-base fine-tune step inside a cudaProfilerStart/Stop window (for `ncu --profile-from-start off`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkg_analogy_b200.config import BASE, LARGE  # noqa: E402
from mkg_analogy_b200.data import MARS_VOCAB, make_finetune_batch  # noqa: E402
from mkg_analogy_b200.trainer import MartTrainer, init_weights_  # noqa: E402

cfg = LARGE if "--large" in sys.argv else BASE
B = 16 if "--large" in sys.argv else 32
steps = int(os.environ.get("PROFILE_STEPS", "1"))
tr = MartTrainer(cfg, MARS_VOCAB, total_steps=1000)
init_weights_(tr.engine, 0)
batch = {k: v.cuda() for k, v in make_finetune_batch(B, 64, cfg.image_size, MARS_VOCAB, seed=1).items()}
for _ in range(3):
    tr.train_step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(steps):
    loss = tr.train_step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", loss.item())
