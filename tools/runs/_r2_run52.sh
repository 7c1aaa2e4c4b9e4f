#!/bin/bash
# round 2, run 52: K1 with atomic Steiner marks: parity, racecheck, speed
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "both_kernels or generic or scaled or below_the_fp32 or synthetic_batch" 2>&1 | tail -3
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -x -q -k "test_infer_cases_both_kernels and (dmv or imdb3) or packed_wire_format" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|Race" | sort | uniq -c | head
timeout 200 python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0,'tests')
import golden_util as G
from bayescard_b200 import _lib as L
from bayescard_b200.engine import DeviceModel
for name in ('census','dmv'):
    tm=G.model(name); dm=DeviceModel(tm, device=0, specialize=False)
    n=1<<20
    st=torch.cuda.current_stream().cuda_stream
    rs=dm.desc_stride(L.DESC_RANGE_U8)
    r=torch.empty((n,rs),dtype=torch.uint8,device='cuda'); out=torch.empty(n,dtype=torch.float32,device='cuda')
    dm.gen_range_queries_device(1,0,n,1,14,r.data_ptr(),st)
    b=torch.empty((n,dm.desc_stride(L.DESC_BITS)),dtype=torch.uint8,device='cuda')
    dm.convert_device(r.data_ptr(), L.DESC_RANGE_U8, b.data_ptr(), L.DESC_BITS, n, st)
    for fmt,buf in ((L.DESC_BITS,b),(L.DESC_RANGE_U8,r)):
        for _ in range(2): dm.run_device(buf.data_ptr(), n, fmt, out.data_ptr(), kernel=L.KERNEL_GENERIC, stream=st)
        torch.cuda.synchronize(); t=time.perf_counter()
        for _ in range(5): dm.run_device(buf.data_ptr(), n, fmt, out.data_ptr(), kernel=L.KERNEL_GENERIC, stream=st)
        torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
        print(name, 'fmt', fmt, 'K1 q/s', round(n/dt/1e6,1),'M')
    dm.close()
PY
