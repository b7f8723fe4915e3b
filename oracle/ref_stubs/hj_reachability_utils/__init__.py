"""Functional stand-in for the un-vendored git dependency ChoiJangho/hj_reachability_utils."""
