python -m pytest tests -q -m gpu -x 2>&1 | tail -12
python bench.py --steps 10 --warmup 3 --cg 200 > gpurun_out/r01_final_c2.json 2> gpurun_out/r01_final_c2.err; cut -c1-200 gpurun_out/r01_final_c2.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 30 --csv --log-file gpurun_out/r01_launches_final_hexp2poisson_n128.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
tools/ncu_capture.sh r01_final_mma_hexp2poisson assemble_gram_mma --grid 64
tools/ncu_capture.sh r01_final_team_hexp2elast assemble_gram_team --phys elasticity --grid 48
tools/ncu_capture.sh r01_final_team_hexp4poisson assemble_gram_team --p 4 --grid 32
tools/ncu_capture.sh r01_final_team_tetp2elast assemble_gram_team --phys elasticity --topo tet --grid 48
du -sh gpurun_out
