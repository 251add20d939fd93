"""Operator API of the AFCM generator hot path, same module names and call signatures as the
reference's models/networks/stylegan3/torch_utils/ops, backed by libafcm_b200.so (sm_100a)."""
