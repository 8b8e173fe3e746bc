import sys, os, torch
sys.path.insert(0, os.getcwd())
from ddnerf_b200 import mlp_tc
from ddnerf_b200.models import base_architectures as BA
net = BA.DepthMipNeRFModel(hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True).to("cuda")
st = mlp_tc._state(net)
for _ in range(3): st.dirty = True; st.refresh()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    st.dirty = True; st.refresh()
e1.record(); torch.cuda.synchronize()
print("pack (weights + bias) us per call:", e0.elapsed_time(e1) * 1000 / 50)
