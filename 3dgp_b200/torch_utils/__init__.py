"""custom_ops.get_plugin shim + op wrappers over lib3dgp_b200.so."""
