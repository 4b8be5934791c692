import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sofima_b200 import flow_field as ff, _native
from oracle import flow_oracle as fo
g = np.load('tests/golden/flow_golden.npz')
pre, post = g['tex160_b4_pre'], g['tex160_b4_post']
starts = np.array([[0,0],[0,40],[0,80],[0,120]], np.int32)
ctx = _native.Context.get()
pre_d = ff._device_image(pre, ctx); post_d = ff._device_image(post, ctx)
p, pre_d, post_d = ff._params(2, pre_d, post_d, None, None, (160,160), (160,160), None, 2, 0.5, 5)
st = torch.from_numpy(starts).cuda()
out = torch.empty((4, 319, 319), device='cuda')
ctx.bind_stream()
rc = _native.lib().sofima_xcorr_images(ctx.handle, ctypes.byref(p), pre_d.data_ptr(), post_d.data_ptr(), None, None, st.data_ptr(), st.data_ptr(), 4, out.data_ptr())
_native.check(ctx.handle, rc)
got = out.cpu().numpy()
center, want = fo.batched_xcorr(pre, post, None, None, (160,160), starts, None)
for b in range(4):
  for b2 in range(4):
    print(b, b2, np.abs(got[b]-want[b2]).max() / np.abs(want[b2]).max())
pk = ff.batched_xcorr_peaks(pre, post, None, None, (160,160), starts, None)
print(pk)
print(fo.batched_xcorr_peaks(pre, post, None, None, (160,160), starts, None))
# peaks on oracle images through the CUDA peak kernels
print(ff._batched_peaks(want, center, 2, 0.5, 5))
