"""d2r_b200 -- B200-native (sm_100a) implementation of D2R's dual-branch dynamic-routing
interaction stack behind the reference's nn.Module API.

Sub-modules
  build        nvcc recipe for csrc/ -> csrc/libd2r_b200.so
  _lib         ctypes binding of include/d2r_b200.h (raises if the library is missing)
  kernels      tensor-level wrappers of the C ABI
  functional   autograd Functions composed of those kernels
  interaction  drop-in mirror of the reference's models/{InteractionModule,DynamicInteraction,
               Cells,Router,SelfAttention,Refinement,XModules}.py
  dp           data-parallel gradient all-reduce plan (NCCL over NVLink)
"""
__version__ = "0.1.0"
