for b in 524288 2097152 4194304 8388608; do for rec in host deferred; do
n=$((1073741824 / b))
./robigo-luculenta_b200/rl_replay --width 1024 --height 1024 --threads 16 --batches $n --batch $b --mode strict --scene 2 --out /tmp/rr --records $rec | python -c "
import sys, json; r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch', r['batch'], r['records'], round(r['mrays_per_s'],1), 'h2d MiB', r['h2d_bytes']>>20, 'd2h MiB', r['d2h_bytes']>>20, 'sec', round(r['seconds'],3))"
done; done
