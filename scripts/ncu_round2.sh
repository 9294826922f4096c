# ncu --set full captures of the round-2 hot kernels (1 GPU).  Reports land in gpurun_out/ncu/, summarised by scripts/summarize_ncu.py.
mkdir -p gpurun_out/ncu
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:conv_tma -s 8 -c 1 -o gpurun_out/ncu/conv_tma_fprop -f python scripts/ncu_fused_step.py --G 8 > gpurun_out/ncu/log_f.txt 2>&1; echo "F rc=$?"
timeout 300 $N -k regex:conv_tma -s 30 -c 1 -o gpurun_out/ncu/conv_tma_bwd_a -f python scripts/ncu_fused_step.py --G 8 > gpurun_out/ncu/log_a.txt 2>&1; echo "A rc=$?"
timeout 300 $N -k regex:conv_tma -s 31 -c 1 -o gpurun_out/ncu/conv_tma_bwd_b -f python scripts/ncu_fused_step.py --G 8 > gpurun_out/ncu/log_b.txt 2>&1; echo "B rc=$?"
timeout 300 $N -k regex:gbn_fwd -s 6 -c 1 -o gpurun_out/ncu/gbn_fwd -f python scripts/ncu_fused_step.py --G 8 > gpurun_out/ncu/log_g1.txt 2>&1; echo "gbn_fwd rc=$?"
timeout 300 $N -k regex:gbn_bwd -s 12 -c 1 -o gpurun_out/ncu/gbn_bwd -f python scripts/ncu_fused_step.py --G 8 > gpurun_out/ncu/log_g2.txt 2>&1; echo "gbn_bwd rc=$?"
timeout 300 $N -k regex:publish_sum -s 1 -c 1 -o gpurun_out/ncu/publish_sum -f python scripts/profile_round.py fedavg 3 8 > gpurun_out/ncu/log_p.txt 2>&1; echo "publish_sum rc=$?"
timeout 300 $N -k regex:fedavg_fullmesh -s 1 -c 1 -o gpurun_out/ncu/fedavg_fullmesh -f python scripts/profile_round.py fedavg 3 8 > gpurun_out/ncu/log_m.txt 2>&1; echo "fullmesh rc=$?"
# launch list of one fused step (G = 8) and of one whole config-2 round
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu/launches_fused_step_g8.csv python scripts/ncu_fused_step.py --G 8 --steps 1 > /dev/null 2>&1; echo "launches rc=$?"
ls -la gpurun_out/ncu | head -30
