for d in 0 1 2 4 8 3 15; do echo "dbg=$d"; RVB_TC_DBG=$d timeout 100 python tools/policy_tc_check.py 2>&1 | grep -E "N    128|N 131072" | sed 's/max.*//'; done
