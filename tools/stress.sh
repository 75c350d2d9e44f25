# Stress of the default PGS kernel at many sizes / sweep counts: every line must end with status_or 0.
cd ${GRAFT_REPO_ROOT:-.}
run() { echo "== $*: $(EGG_SYNC_DEBUG=1 timeout 600 python tools/profile_run.py "$@" 2>&1 | tail -n 1 | cut -c 1-140)"; }
for rep in 1 2; do
run c3 4096 5 2
run c3 4096 20 2
run c3 16384 5 2
run c3 16384 20 2
run c3 65536 5 1
run c2 4096 50 2
run c2 65536 50 2
run c5 8192 50 2
run c5 131072 50 2
done
