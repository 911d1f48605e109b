"""Print the headline numbers of a bench.py JSON line: bench_line.py <label> <file>"""
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1], f'{d["value"]:.2f} img/s  e2e {d["e2e"]["value"]:.2f}  conv frac {r["frac"]:.3f}  conv {r["conv_ms_per_step"] * 1e3:.0f} us  '
          f'elementwise {r["elementwise_ms_per_step"] * 1e3:.0f} us  graph {r["graph_ms"]:.3f} ms')
except Exception as e:  # noqa: BLE001
    print(sys.argv[1], "ERR", e)
