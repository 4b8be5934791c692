"""sofima_b200: B200-native (sm_100a) backend for SOFIMA's two hot paths.

  sofima_b200.flow_field   drop-in for sofima.flow_field  (patch-flow estimator)
  sofima_b200.mesh         drop-in for sofima.mesh        (spring-mesh relaxation)

Both modules keep the reference's public names and signatures and run on
hand-written CUDA behind the C ABI of include/sofima_b200.h.  There is no CPU
fallback: without the built library and a B200 the calls raise.
"""

__version__ = '0.1.0'
