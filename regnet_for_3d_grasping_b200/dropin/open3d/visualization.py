def draw_geometries(*args, **kwargs):
    raise RuntimeError("open3d stand-in: no visualisation back end (install the real open3d to draw geometries)")
