"""One clic_find_correspondence call at the production shape (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clic_b200 import association as A, estimator as E, synthetic as S
sc = S.make_feature_scene(n_map_surf=50000, n_map_corner=8000, n_surf=1500, n_corner=1500, seed=3)
fm = A.FeatureMap(sc.map_surf, sc.map_corner, E.cuda_library())
for _ in range(3):
    c = fm.FindCorrespondence(sc.corner_L, sc.corner_t, sc.corner_M, sc.surf_L, sc.surf_t, sc.surf_M)
print(len(c.line_t), len(c.plane_t), fm.launches())
fm.close()
