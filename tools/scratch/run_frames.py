import numpy as np, sys
sys.path.insert(0, '/root/repo')
from multicameratracking_b200 import synth
from multicameratracking_b200 import host as mhost
n_frames, n_obj = 8, 8
over = dict(feature_type=3, weak_type=1, n_particles=2000, n_haar=250)
seq = synth.Sequence(640, 480, n_obj, n_frames=n_frames)
p = dict(mhost.DEFAULTS); p.update(over)
cam = mhost.Camera(0, n_frames)
cam.set_frame(*seq.frame(0))
for o in range(n_obj):
    cam.add_object(seq.box(0, o), rng_seed=o + 1, **over)
cam.init_trackers()
N = p["n_particles"]
for t in range(1, n_frames):
    cam.set_frame(*seq.frame(t))
    nz = np.stack([synth.noise(N, (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), obj=o, t=t) for o in range(n_obj)])
    cam.track(t, nz, 3)
cam.sync()
print("done")
