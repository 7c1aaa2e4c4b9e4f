import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, golden_util as G
from bayescard_b200 import _lib as L
from bayescard_b200.engine import DeviceModel
from bayescard_b200.decode import PredicateCompiler
for name in ("census","imdb1"):
    m=G.model(name); dm=DeviceModel(m, device=0, specialize=True)
    host=dm.gen_range_queries_host(1,0,64,1,5)
    pc=PredicateCompiler(m)
    from bayescard_b200.decode import unpack_ranges
    lo,hi=unpack_ranges(m,host); bits=pc.pack_bits(lo,hi)
    for kname,kernel in (('auto',L.KERNEL_AUTO),('spec',L.KERNEL_SPEC),('fused',L.KERNEL_FUSED),('generic',L.KERNEL_GENERIC)):
      for fmt,desc in ((L.DESC_RANGE_U8,host),(L.DESC_BITS,bits)):
        try:
            for _ in range(50): dm.run_host(desc[:1], fmt, None, kernel)
        except Exception as e:
            print(name, kname, fmt, 'n/a'); continue
        ts=[]
        for i in range(300):
            t=time.perf_counter(); dm.run_host(desc[i%64:i%64+1], fmt, None, kernel); ts.append(time.perf_counter()-t)
        print(name, kname, fmt, 'run_host B=1 p50 us', round(np.median(ts)*1e6,1), 'p99', round(np.percentile(ts,99)*1e6,1))
    dm.close()
