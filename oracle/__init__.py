"""ORACLE — test infrastructure only.

CPU restatement (C++17, no third-party dependencies) of the reference's BA and RANSAC hot
paths. Imported ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, as the checker and the CPU baseline. pytheiasfm_b200/ never imports it.
"""
