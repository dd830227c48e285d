"""Python wrappers of the sm_100a kernels behind the op names the reference modules import (bias_act, upfirdn2d, conv2d_gradfix, ...)."""
