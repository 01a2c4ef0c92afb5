O=gpurun_out
mkdir -p $O
timeout 100 python profiles/prof_h7_phases.py 64 | tee $O/h7_fwd_phases_b64.txt | head -8
timeout 100 python profiles/prof_h7_phases.py 2048 | tee $O/h7_fwd_phases_b2048.txt | head -24
timeout 100 python profiles/time_h7_all.py 64 2>&1 | tail -8 | cut -c1-300
