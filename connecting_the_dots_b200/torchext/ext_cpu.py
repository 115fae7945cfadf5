"""Placeholder for the reference's `ext_cpu` module (torchext/ext/ext_cpu.cpp:189-198).

The reference imports ext_cpu unconditionally (torchext/functions.py:2) and routes CPU tensors to it.
This framework has no CPU compute path by design: every entry point raises.  Callers with host
buffers use the C ABI's ctd_host_* functions (include/ctd_b200.h), which run on the GPU."""


def _no_cpu(name):
    def fn(*args, **kwargs):
        raise RuntimeError("torchext.%s: connecting_the_dots_b200 has no CPU implementation; move the tensors to a "
                           "CUDA device (or use the ctd_host_* C entry points)" % name)
    fn.__name__ = name
    return fn


nn_cpu = _no_cpu("nn_cpu")
crosscheck_cpu = _no_cpu("crosscheck_cpu")
proj_nn_cpu = _no_cpu("proj_nn_cpu")
xcorrvol_cpu = _no_cpu("xcorrvol_cpu")
photometric_loss_forward = _no_cpu("photometric_loss_forward")
photometric_loss_backward = _no_cpu("photometric_loss_backward")
