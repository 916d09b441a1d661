"""Matcher-related configuration keys.

When the real ``s2p`` package is importable its global ``cfg`` dict is used, so a
``config.json`` overlay keeps working unchanged.  Otherwise a dict with the same keys and the
reference's default values (s2p/config.py:18,43-46,77,136-160) stands in.
"""
try:  # pragma: no cover - depends on the installation
    from s2p.config import cfg  # noqa: F401
except Exception:
    cfg = {
        "temporary_dir": "s2p_tmp",
        "omp_num_threads": 1,
        "max_processes_stereo_matching": None,
        "max_disp_range": None,
        "matching_algorithm": "mgm",
        "census_ncc_win": 5,
        "stereo_speckle_filter": 25,
        "stereo_regularity_multiplier": 1.0,
        "mgm_nb_directions": 8,
        "mgm_timeout": 600,
        "mgm_leftright_threshold": 1.0,
        "mgm_leftright_control": 1,
        "mgm_mindiff_control": -1,
    }
