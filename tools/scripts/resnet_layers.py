import sys
sys.path.insert(0, '.')
import numpy as np
from fyusenet_b200 import hostapi
rng = np.random.default_rng(0)
for batch in (1, 32, 128):
    net = hostapi.ResNet50(batch=batch)
    net.load_weights((rng.standard_normal(25576046) * 0.01).astype(np.float32))
    net.setup()
    net.set_input(rng.random((batch, 224, 224, 3), dtype=np.float32))
    net.forward()
    net.enable_timings(True)
    for _ in range(3):
        net.forward()
    net.finish()
    rows = []
    for l in net.layers():
        ms = net.layer_timing(l["number"])[0] / 3
        rows.append((ms, l["name"], l.get("channels"), l.get("width"), l.get("height"), l.get("family")))
    tot = sum(r[0] for r in rows)
    print(f"batch {batch}: total {tot:.3f} ms")
    kinds = {}
    for r in rows:
        k = r[1].rstrip('0123456789')
        kinds[k] = kinds.get(k, (0, 0.0))
        kinds[k] = (kinds[k][0] + 1, kinds[k][1] + r[0])
    print('  by layer type: ' + ', '.join(f'{k} x{n}: {t * 1e3:.0f} us' for k, (n, t) in sorted(kinds.items(), key=lambda kv: -kv[1][1])))
    for r in sorted(rows, reverse=True)[:24]:
        print(f"  {r[0]*1e3:8.1f} us  {r[1]:12s} out {r[2]}ch {r[3]}x{r[4]} fam {r[5]}")
    net.destroy()
