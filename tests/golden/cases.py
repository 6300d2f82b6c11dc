"""Model specifications of the golden filter cases (shared by make_golden.py and the tests)."""
CASES = {
    # name: (latents [[(kind, ls, var), ...], ...], full_state_obs, T, nan_frac, seed)
    "m32": ([[("m32", 1.0, 1.3)]], False, 40, 0.1, 1),
    "m52": ([[("m52", 0.7, 0.9)]], False, 40, 0.1, 2),
    "m72": ([[("m72", 1.2, 1.1)]], False, 40, 0.1, 3),
    "sum_m32_m52": ([[("m32", 1.0, 1.3), ("m52", 0.4, 0.5)]], False, 33, 0.1, 4),
    "indep_m32x2": ([[("m32", 1.0, 1.3)], [("m32", 0.4, 0.5)]], False, 33, 0.2, 5),
    "m52_fullstate": ([[("m52", 0.7, 0.9)]], True, 33, 0.2, 6),
    "indep_m32_m52_fullstate": ([[("m32", 1.0, 1.3)], [("m52", 0.6, 0.8)]], True, 25, 0.25, 7),
}


