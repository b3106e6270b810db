import os, sys
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from massivedatans_b200 import synth, _lib
from massivedatans_b200.likelihood import ResidentDataset
n, nx, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
chk = tuple(int(v) for v in sys.argv[4].split(','))      # variant under test
pol = tuple(int(v) for v in sys.argv[5].split(','))      # variant that pollutes the workspace
x, y, _ = synth.realistic(n, nx=nx) if nx == 1000 else synth.horns(n, nx=nx, legacy=False, seed=1000)
ds = ResidentDataset(x, y)
mask = synth.masks(n, seed=3)[os.environ.get('DBG_MASK', 'all')]
n_act = int(mask.sum())
ds.set_mask(None if mask.all() else mask)
pts = synth.parameter_points(K, seed=7)
ds.stage_params(pts)
def run(tun, reps=1):
    ds.set_tuning(*tun)
    o = numpy.empty((K, n_act))
    for _ in range(reps):
        ds.launch_clike(0.01, -0.5)
    ds.fetch(o)
    return o
ref = run((3, 1, 0, 0))
nbad = 0
for it in range(40):
    run(pol, reps=1 + it % 5)
    o = run(chk, reps=1 + (it // 5) % 3)
    rel = numpy.abs(o - ref) / numpy.abs(ref)
    bad = numpy.argwhere(rel > 1e-9)
    if len(bad):
        nbad += 1
print(sys.argv[1:], 'graph' if not os.environ.get('MDNS_NO_GRAPH') else 'nograph', 'bad iterations', nbad, 'of 40', _lib.load().mdns_last_kernel())
