#!/bin/bash
cd "$GRAFT_REPO_ROOT"
RF_KNN_COOP=1 timeout 1200 python -m pytest tests -x -q -m gpu -k "knn" > gpurun_out/r2s2_pytest_knn.log 2>&1; echo "pytest(coop=0) rc=$?"; tail -3 gpurun_out/r2s2_pytest_knn.log
for coop in 1; do
for b in encoded random; do
RF_KNN_COOP=$coop timeout 600 python bench.py --workload retrieval --bank $b --no-cpu-baseline --steps 3 > /tmp/b.json 2>/dev/null
python -c "
import json
l=json.load(open('/tmp/b.json')); print('coop=$coop $b', round(l['value']), l['breakdown_ms']['knn'], l.get('knn_stats'))"
done
RF_KNN_COOP=$coop timeout 900 python bench.py --workload sweep --steps 2 --warmup 1 > /tmp/s.json 2> /dev/null
python -c "
import json
l=json.load(open('/tmp/s.json'))
for s in l['sweep']: print('coop=$coop', {k:s[k] for k in ('k','knn_bulk_ms','knn_64chunks_ms')})"
done
