python - <<'PY'
import torch, json, sys
sys.path.insert(0, ".")
from road_segmentation_unet_b200 import unet, ops, _lib
net = unet.UNet(6, 64, True, 32, 764)
x = torch.rand(32, 764, 764, 3, device="cuda"); y = (torch.rand(32, 388, 388, device="cuda") < 0.3).to(torch.uint8)
for _ in range(2): net.train_step(x, y)
torch.cuda.synchronize()
_lib.load().rsu_reset_launch_count()
ops.profile_start()
net.train_step(x, y)
rec = ops.profile_stop()
print(json.dumps(_lib.launch_histogram()))
per = {}
for kind, layer, fl, ms in rec:
    a = per.setdefault(kind + "|" + str(layer), [0.0, 0.0]); a[0] += fl; a[1] += ms
for k, (fl, ms) in sorted(per.items(), key=lambda kv: -kv[1][1])[:16]:
    print("%-45s %7.3f ms %6.0f TF/s" % (k, ms, fl / ms / 1e9))
PY
