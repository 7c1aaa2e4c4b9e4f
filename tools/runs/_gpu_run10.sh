mkdir -p gpurun_out
b() { env "$@" timeout 120 python bench.py --model $M --steps 10 --cpu-seconds 1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(r['value']/1e9, 4), 'Gq/s frac', round(r['roofline']['frac'], 4), 'relerr', r['rel_err_max_vs_fp64_oracle'])
"; }
{
M=imdb1
for v in X=1 "BC_SPEC_THREADS=384 BC_SPEC_MIN_BLOCKS=1" "BC_SPEC_THREADS=384 BC_SPEC_MIN_BLOCKS=1 BC_SPEC_SYNC_EVERY=256" "BC_SPEC_THREADS=384 BC_SPEC_MIN_BLOCKS=1 BC_SPEC_SYNC_EVERY=1024" "BC_SPEC_THREADS=256 BC_SPEC_MIN_BLOCKS=1 BC_SPEC_SYNC_EVERY=512" BC_SPEC_SYNC_EVERY=512; do echo "== imdb1 $v"; b $v; done
M=dmv
for v in "BC_SPEC_THREADS=256 BC_SPEC_MIN_BLOCKS=1 BC_SPEC_SYNC_EVERY=512" BC_SPEC_SYNC_EVERY=512; do echo "== dmv $v"; b $v; done
} > gpurun_out/s10_spec_align.txt 2>&1
cat gpurun_out/s10_spec_align.txt
