#!/bin/bash
# GMWB timing at one size for the three driver depths (0: iterations one after the other, 1: control
# search overlapped with the line pass, 2: + speculative trailing pass).  usage: tools/profile_gmwb_depth.sh refine a_points timesteps0
for d in 0 1 2; do
  echo "QPDE_GMWB_DEPTH=$d"; QPDE_GMWB_DEPTH=$d python tools/profile_gmwb.py "$@"
done
