"""Constants of the reference's global_variables.py that the hot path reads (global_variables.py:15-21)."""
g_zero_tol = 1.0e-6
small_area_threshold = 0.02
EXTRUSION_OPERATION_DICT = {"NewBodyFeatureOperation": 0, "JoinFeatureOperation": 0, "CutFeatureOperation": 1,
                            "IntersectFeatureOperation": 2}
