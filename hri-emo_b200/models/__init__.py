"""Drop-in replacements for the reference's ``models`` package (same module paths,
class names, constructor arguments, forward signatures and state_dict keys), with
the forward pass executed by the sm_100a kernels of libhriemo_b200.so."""
