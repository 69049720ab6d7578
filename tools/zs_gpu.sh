#!/bin/bash
# GPU loop for the residual coder: real C2 parts one by one with the phase profile (build with EXTRA=-DZE_PROF for the breakdown)
AGCGPU_TRACE=1 timeout 300 python tools/zs_parts_prof.py 2>&1 | grep -E "^x|frame 0 \((7502|107362|1350047)|counts|packs|rror" | awk '!seen[$0]++' | head -30
