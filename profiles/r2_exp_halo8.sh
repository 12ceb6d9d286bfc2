N=${1:-8}
run() { # name, env...
  name=$1; shift
  env "$@" VB_TRACE=2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu > gpurun_out/exp${N}_$name.json 2> gpurun_out/exp${N}_$name.err
  echo "== $name: $(python -c "import json; d=json.loads(open('gpurun_out/exp${N}_$name.json').read().strip().splitlines()[-1]); print(d['ms_per_step'])")"
  grep "vb mark\|vb halo" gpurun_out/exp${N}_$name.err | tail -$((N*26)) | grep "sweep of\|push of\|last mark" | sed 's/: [0-9]* entries//; s/: [0-9]* B//; s/(at.*//' | awk '{k=$3" "$4" "$5" "$6" "$7; v=$(NF-1); s[k]+=v; c[k]++; if(v>m[k])m[k]=v} END{for(k in s) printf "   %-40s mean %.3f max %.3f (n=%d)\n", k, s[k]/c[k], m[k], c[k]}' | sort
  grep "last mark" gpurun_out/exp${N}_$name.err | tail -$N | awk '{print $(NF-1)}' | sort -n | tail -1 | sed 's/^/   main stream ends at (max) /'
}
run ce_overlap VB_HALO_CE=1
run push_overlap VB_HALO_CE=0
run push_nooverlap VB_HALO_OVERLAP=0 VB_HALO_CE=0
