"""One eager train step (cfg 3 shapes) between cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ifseg_b200.segofa import SegOFAModel  # noqa: E402
from ifseg_b200.synthetic import generate_state_dict, synthetic_train_sample  # noqa: E402
from ifseg_b200.trainer import SegOFATrainer  # noqa: E402

arch, size, nseg, batch, _ = bench.CONFIGS[3]
model = SegOFAModel.from_config(arch, nseg, size)
model.load_state_dict(generate_state_dict(model.cfg, 0), strict=True)
model = model.cuda()
tr = SegOFATrainer(model)
smp = tr.to_device(synthetic_train_sample(model.cfg, batch, size, seed=1, src_tokens=bench.prompt_tokens(nseg)))
for _ in range(2):
    tr.train_step(smp, check_pads=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.train_step(smp, check_pads=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
