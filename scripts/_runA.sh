set -x
mkdir -p gpurun_out
python scripts/prof_cg_ops.py > gpurun_out/r2_cgops_plain.log 2>&1; echo rc=$?
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r2_cg_ops python scripts/prof_cg_ops.py > gpurun_out/r2_ncu_cgops.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_factor_kernel -s 24 -c 3 -o gpurun_out/r2_panel_factor python scripts/prof_chol.py 5000 1 > gpurun_out/r2_ncu_panel.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma_bulk -s 40 -c 2 -o gpurun_out/r2_chol_bulk python scripts/prof_chol.py 20000 1 > gpurun_out/r2_ncu_cholbulk.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
