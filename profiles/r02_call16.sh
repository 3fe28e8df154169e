export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
FULL="ncu --set full --clock-control none --import-source on"
SLC_DYNA_WARP=1 $FULL -k regex:"dyna_warp_kernel" -s 1 -c 1 -o gpurun_out/r02p_prof_dyn_warp -f python bench.py --path dynamic --steps 1 --warmup 1 --dyna-frames 30 > gpurun_out/r02p_ncu_warp.log 2>&1
$FULL -k regex:"dyna_fused_kernel" -s 1 -c 1 -o gpurun_out/r02p_prof_dyn_rowmap -f python bench.py --path dynamic --steps 1 --warmup 1 --dyna-frames 30 > gpurun_out/r02p_ncu_rowmap.log 2>&1
ls -la gpurun_out/r02p_*
