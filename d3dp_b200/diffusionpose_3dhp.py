"""MPI-INF-3DHP variant of `D3DP` (reference: common/diffusionpose_3dhp.py — identical to common/diffusionpose.py
except that poses are exchanged in millimetres: sampler outputs are multiplied by 1000 (:212,256,287) and the
training input is divided by 1000 (:280)).  The x1000 is applied by the DDIM kernel when it stores the per-step
predictions (`d3dp_config.output_scale`), not by an extra pass over the output."""
from .diffusionpose import D3DP as _D3DP

__all__ = ["D3DP"]


class D3DP(_D3DP):
    OUTPUT_SCALE = 1000.0
