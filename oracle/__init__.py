"""CPU oracle for the MVSNet depth-inference hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (scene_3dreconstruction_mvsnet_b200) never does.
"""
