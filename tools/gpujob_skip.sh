for f in 1 2 3; do echo "=== skip flags $f"; GPUCHAN_DEBUG_SKIP=$f python profiles/tools/tc_role_stamps.py 2>&1 | awk '/^mma/{p=1} p' | sed -n 1,9p; done
