"""Runtime flags, same names as the reference's ``dflex/config.py:10-12``."""
no_grad = False      # True: integrator updates the state in place and keeps no tape (sim.py:2201-2207)
check_grad = False   # accepted for compatibility; gradient checks live in tests/ here
verify_fp = False    # True: assert finite outputs after every integrator call
