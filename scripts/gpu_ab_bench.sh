cp sp-gan_b200/libspgan_b200.so /tmp/lib_current.so
for v in prev cur prev cur; do
  cp sp-gan_b200/variants/$v.so sp-gan_b200/libspgan_b200.so
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_share'].get('gemm'), d['kernel_share'].get('gemm_wgrad_fused'))"
done
cp /tmp/lib_current.so sp-gan_b200/libspgan_b200.so
