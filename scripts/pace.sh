for P in 0 100 200 400 800 1600; do echo "== ZG_PACE=$P"; ZG_PACE=$P timeout 200 python scripts/phase_profile.py 124M 32 2>&1 | grep -E "unprofiled|sum"; done
