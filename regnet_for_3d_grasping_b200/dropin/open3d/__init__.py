"""Stand-in for the parts of `open3d` (0.7-0.9 API) that the reference touches outside its visualisation scripts:

  geometry.PointCloud (points / colors / normals, transform, voxel_down_sample, estimate_normals, normalize_normals,
  orient_normals_towards_camera_location), geometry.KDTreeSearchParam{KNN,Radius,Hybrid}, geometry.KDTreeFlann
  (search_knn_vector_3d / search_radius_vector_3d / search_hybrid_vector_3d), utility.Vector3dVector,
  io.read_point_cloud / write_point_cloud for PCD files (ascii and binary, x y z [rgb | rgba]).

Call sites: test.py:102-106, dataset_utils/eval_score/eval_utils/pointcloud.py:8-43, torch_scene_point_cloud.py:10-25,
evaluation_data_generator.py:59-61,247-261.  numpy + scipy.spatial.cKDTree underneath; `visualization` only says that no
display is available.  It exists so that `import open3d` -- which the reference does at module import time in files that
training needs -- succeeds without the real package; it is not part of the accelerated path."""
from _bootstrap import installed_elsewhere as _installed_elsewhere   # the drop-in directory is on sys.path (that is how this package was found)

if _installed_elsewhere("open3d") is None:      # otherwise sys.modules["open3d"] now is the real package
    from . import geometry, io, utility, visualization  # noqa: F401

    __version__ = "0.0-regnet-b200-standin"
