"""Physical constants of the pulse-DDM task (same names and values as the reference's
constants.py:1-5, which the hot path reads at call time)."""
DT = 1e-6            # unused by the reference pipeline; kept for the dt=1e-6 stress schedule
DT_CHOICE = 5e-4     # Euler step of the RT/choice simulator [s]
T_MAX = 8.0          # trial length [s]
PULSE_INTERVAL = 0.1 # time between pulses [s]
