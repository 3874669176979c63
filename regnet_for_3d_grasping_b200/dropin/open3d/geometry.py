import numpy as np
from scipy.spatial import cKDTree


class KDTreeSearchParamKNN:
    def __init__(self, knn=30):
        self.knn = int(knn)


class KDTreeSearchParamRadius:
    def __init__(self, radius):
        self.radius = float(radius)


class KDTreeSearchParamHybrid:
    def __init__(self, radius, max_nn):
        self.radius, self.max_nn = float(radius), int(max_nn)


class PointCloud:
    def __init__(self):
        self._points = np.zeros((0, 3))
        self._colors = np.zeros((0, 3))
        self._normals = np.zeros((0, 3))

    points = property(lambda self: self._points, lambda self, v: setattr(self, "_points", np.asarray(v, dtype=np.float64).reshape(-1, 3)))
    colors = property(lambda self: self._colors, lambda self, v: setattr(self, "_colors", np.asarray(v, dtype=np.float64).reshape(-1, 3)))
    normals = property(lambda self: self._normals, lambda self, v: setattr(self, "_normals", np.asarray(v, dtype=np.float64).reshape(-1, 3)))

    def has_points(self):
        return len(self._points) > 0

    def has_colors(self):
        return len(self._colors) == len(self._points) > 0

    def has_normals(self):
        return len(self._normals) == len(self._points) > 0

    def transform(self, T):
        T = np.asarray(T, dtype=np.float64)
        self._points = self._points @ T[:3, :3].T + T[:3, 3]
        if self.has_normals():
            self._normals = self._normals @ T[:3, :3].T
        return self

    def voxel_down_sample(self, voxel_size):
        """One point per occupied voxel: the mean of the voxel's points (colours / normals averaged alike)."""
        out = PointCloud()
        if not self.has_points():
            return out
        key = np.floor((self._points - self._points.min(0)) / float(voxel_size)).astype(np.int64)
        _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
        inv = inv.reshape(-1)

        def mean(a):
            acc = np.zeros((len(cnt), 3))
            np.add.at(acc, inv, a)
            return acc / cnt[:, None]

        out.points = mean(self._points)
        if self.has_colors():
            out.colors = mean(self._colors)
        if self.has_normals():
            out.normals = mean(self._normals)
        return out

    def estimate_normals(self, search_param=None, fast_normal_computation=True):
        """Normal = eigenvector of the smallest eigenvalue of the neighbourhood covariance (sign arbitrary until
        orient_normals_* is called), neighbourhoods by kNN, radius or both."""
        sp = search_param or KDTreeSearchParamKNN()
        pts = self._points
        tree = cKDTree(pts)
        normals = np.tile(np.array([0.0, 0.0, 1.0]), (len(pts), 1))
        if isinstance(sp, KDTreeSearchParamRadius):
            nbrs = tree.query_ball_point(pts, sp.radius)
        else:
            k = min(sp.knn if isinstance(sp, KDTreeSearchParamKNN) else sp.max_nn, len(pts))
            bound = np.inf if isinstance(sp, KDTreeSearchParamKNN) else sp.radius
            dist, idx = tree.query(pts, k=k, distance_upper_bound=bound)
            idx = np.asarray(idx).reshape(len(pts), -1)
            nbrs = [row[row < len(pts)] for row in idx]
        for i, nb in enumerate(nbrs):
            if len(nb) < 3:
                continue
            q = pts[np.asarray(nb)]
            w, v = np.linalg.eigh(np.cov((q - q.mean(0)).T))
            normals[i] = v[:, 0]
        self._normals = normals
        return True

    def normalize_normals(self):
        n = np.linalg.norm(self._normals, axis=1, keepdims=True)
        self._normals = np.divide(self._normals, n, out=self._normals.copy(), where=n > 0)
        return self

    def orient_normals_towards_camera_location(self, camera_location=np.zeros(3)):
        to_cam = np.asarray(camera_location, dtype=np.float64).reshape(1, 3) - self._points
        flip = (self._normals * to_cam).sum(1) < 0
        self._normals[flip] *= -1.0
        return True


class KDTreeFlann:
    """Results follow open3d: (count, indices, squared distances), neighbours in ascending distance."""

    def __init__(self, geometry=None):
        self._tree = None
        if geometry is not None:
            self.set_geometry(geometry)

    def set_geometry(self, geometry):
        pts = geometry.points if isinstance(geometry, PointCloud) else np.asarray(geometry, dtype=np.float64).reshape(-1, 3)
        self._n = len(pts)
        self._tree = cKDTree(pts)
        return True

    def search_knn_vector_3d(self, query, knn):
        d, i = self._tree.query(np.asarray(query, dtype=np.float64).reshape(3), k=min(int(knn), self._n))
        d, i = np.atleast_1d(d), np.atleast_1d(i)
        return len(i), [int(v) for v in i], [float(v) ** 2 for v in d]

    def search_radius_vector_3d(self, query, radius):
        q = np.asarray(query, dtype=np.float64).reshape(3)
        i = np.asarray(self._tree.query_ball_point(q, float(radius)), dtype=np.int64)
        d2 = ((self._tree.data[i] - q) ** 2).sum(1) if len(i) else np.zeros(0)
        order = np.argsort(d2, kind="stable")
        return len(i), [int(v) for v in i[order]], [float(v) for v in d2[order]]

    def search_hybrid_vector_3d(self, query, radius, max_nn):
        k, i, d2 = self.search_radius_vector_3d(query, radius)
        k = min(k, int(max_nn))
        return k, i[:k], d2[:k]
