# sweep-kernel shape variants (built into vahana.jl_b200/csrc/build/variants/, selected with VAHANA_B200_LIB): kernel ms of the HK-100M step
for v in A B C; do
  VAHANA_B200_LIB=$PWD/vahana.jl_b200/csrc/build/variants/libvahana_b200_$v.so timeout 120 python bench.py --steps 5 --no-cpu --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant $v', d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['opinion_sum'])"
done
