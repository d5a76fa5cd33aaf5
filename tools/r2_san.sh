#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-san}; mkdir -p $O
TCR_RESIDENT_VERBOSE=1 python -c "
import sys; sys.path.insert(0,'.')
import tcresnet_b200
from tcresnet_b200.engine import Engine
e=Engine(max_batch=512)
import torch
p,s,m=e.new_variables(0)
wav=torch.rand(512,16000,device='cuda')*2-1
hot=torch.nn.functional.one_hot(torch.randint(0,12,(512,),device='cuda'),12).float()
e.train_step(wav,hot,p,s,m,0.1); torch.cuda.synchronize(); print('ok 512')
" > $O/sizes.txt 2>&1; tail -14 $O/sizes.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from tcr_harness import TorchBackend
from parity_cases import run_case
b = TorchBackend()
print(run_case(b, model='TCResNet8', wm=1.0, n=1))
" > $O/sanitizer.txt 2>&1 ; echo "sanitizer rc=$?"
grep -m1 -B2 -A25 "Invalid\|Error:" $O/sanitizer.txt | head -60
