"""Debug aid: compares generate_mesh with scipy's Delaunay on the same vertices and reports, for every
triangle that only one side has, how deep the worst point lies inside its circumcircle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from superscreen_b200 import meshgen
from superscreen_b200.geometry import circle
from test_gpu_meshgen import _scipy_region_triangles, _canon

outer, hole = circle(4.0, 100), circle(2.0, 60)
points, triangles = meshgen.generate_mesh(outer, hole_coords=[hole], min_points=3000, max_edge_length=0.3)
ref = _scipy_region_triangles(points, [outer, hole])
a, b = _canon(triangles), _canon(ref)

def violation(t):
    p = points[list(t)]
    ax, ay, bx, by, cx, cy = *p[0], *p[1], *p[2]
    d = 2 * (ax * (by - cy) + bx * (cy - ay) + cx * (ay - by))
    ux = ((ax**2 + ay**2) * (by - cy) + (bx**2 + by**2) * (cy - ay) + (cx**2 + cy**2) * (ay - by)) / d
    uy = ((ax**2 + ay**2) * (cx - bx) + (bx**2 + by**2) * (ax - cx) + (cx**2 + cy**2) * (bx - ax)) / d
    R = np.hypot(ax - ux, ay - uy)
    dist = np.hypot(points[:, 0] - ux, points[:, 1] - uy)
    dist[list(t)] = np.inf
    k = int(np.argmin(dist))
    return (R - dist[k]) / R, k, R

print("mine only:", len(a - b), " scipy only:", len(b - a), " of", len(a))
for name, s in (("mine", a - b), ("scipy", b - a)):
    for t in sorted(s)[:12]:
        v, k, R = violation(t)
        print(f"  {name} {t}: worst point {k} inside by {v:+.3e} R (R = {R:.3f}); vertices {np.round(points[list(t)], 4).tolist()}")
