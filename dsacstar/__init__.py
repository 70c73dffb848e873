"""Drop-in replacement of the reference's `dsacstar` extension module (reference dsacstar/dsacstar.cpp:898-899).

`import dsacstar; dsacstar.forward_rgb(...)` keeps the reference's positional signature; the work runs in the
sm_100a CUDA solver of libacez.so. `forward_rgb_batch` is the batched device-resident entry.
"""
from acezero_b200.dsac import forward_rgb, forward_rgb_batch  # noqa: F401
