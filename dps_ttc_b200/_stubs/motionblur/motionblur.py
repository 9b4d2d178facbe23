"""Stand-in for motionblur.motionblur.Kernel: same constructor and `.kernelMatrix`, kernel from
dps_ttc_b200.tables.motion_kernel (numpy global RNG).  Parity of kernel GENERATION is unpinned."""
from dps_ttc_b200.tables import MotionKernel as Kernel  # noqa: F401
