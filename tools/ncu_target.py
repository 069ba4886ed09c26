#!/usr/bin/env python
"""Short, deterministic target for `ncu`: three eager steps of the bench workload (OV-7B, bf16,
1 video x 64 frames).  Per step the library launches, in order: 25 gemm_tc_kernel, 5 attn_tc_kernel,
9 layernorm_kernel, 1 pool_pe_kernel, 2 add_pe_kernel, 1 assemble_kernel.  Profile the third step:

  ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 50 -c 25 -o gpurun_out/gemm  python tools/ncu_target.py
  ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 10 -c 5  -o gpurun_out/attn  python tools/ncu_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mavlm_b200 import synthetic  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pipe, _ = synthetic.build_pipeline(3584, 1152, dtype=torch.bfloat16, chunk_size=32, device="cuda:0")
x = synthetic.synthetic_tower_tokens(1, 64).to("cuda:0")
idx = torch.arange(64)[None]
for _ in range(steps):
    pipe(x, idx, return_states=False)
torch.cuda.synchronize()
print("ok")
