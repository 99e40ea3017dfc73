"""Physical and numerical constants — same VALUES as the reference's `constants.py:13-29`
(they are physics inputs, so they must match digit for digit)."""
import math

pi = 3.14159265358979323846264338327950288419716939937510
deg2rad = pi / 180.0
rad2deg = 1.0 / deg2rad
rearth = 6.3712e6          # Earth radius [m]
omega = 7.2921e-5          # Earth rotation rate [1/s]

one = 1.0
zero = 0.0

hour = 3600.0
day = 24.0 * hour

delt = 1.0e-8              # |Im(root)| below this counts as a real root (bs.py:1030)
undef = math.nan           # missing / not propagating
