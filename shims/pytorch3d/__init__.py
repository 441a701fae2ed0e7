"""Import-name stand-in for the one pytorch3d entry point tinysplat calls (`pytorch3d.ops.knn_points`).
Opt-in: put `<repo>/shims` on `sys.path`."""
