"""sofima_b200: B200-native (sm_100a) backend for SOFIMA's two hot paths.

  sofima_b200.flow_field   drop-in for sofima.flow_field  (patch-flow estimator)
  sofima_b200.mesh         drop-in for sofima.mesh        (spring-mesh relaxation)

Both modules keep the reference's public names and signatures and run on
hand-written CUDA behind the C ABI of include/sofima_b200.h.  There is no CPU
fallback: without the built library and a B200 the calls raise.
"""

__version__ = '0.1.0'


def release_memory(keep_bytes: int = 0, device=None) -> None:
  """Frees the calling thread's cached device scratch (row-spectra cache, spectra and image
  scratch of the flow path; several GB after a large `flow_field` call) so that other
  users of the GPU -- torch's allocator, the mesh solver -- can have it.  Optional: the
  buffers are otherwise kept for the next call of the same geometry."""
  import sys
  from . import _native
  _native.Context.get(device).trim(keep_bytes)
  warp = sys.modules.get(__name__ + '.warp')
  if warp is not None and not keep_bytes:
    import torch
    torch.cuda.synchronize()
    warp._PINNED.clear()  # pinned section blocks of warp_subvolume
