for pf in ${DS_PFS:-4 12 16}; do
  MM_DS_PF=$pf timeout 200 python scripts/gpu_decode_bench.py > gpurun_out/ds_pf_$pf.json 2> gpurun_out/ds_pf_$pf.err
done
