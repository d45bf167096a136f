"""Physical and astronomical constants of the Qingdai system (values of pygcm/constants.py:9-35)."""

G = 6.67430e-11
SIGMA = 5.670374e-8
M_SUN = 1.989e30
L_SUN = 3.828e26
AU = 1.496e11

M_A = 0.914 * M_SUN
L_A = 0.7 * L_SUN
M_B = 0.8 * M_SUN
L_B = 0.410 * L_SUN
M_TOTAL_STARS = M_A + M_B
A_BINARY = 0.5 * AU

A_PLANET = 1.32 * AU
PLANET_RADIUS = 6.371e6
PLANET_ALBEDO = 0.3
PLANET_OMEGA = 8.726646259971648e-5
PLANET_AXIAL_TILT = 27.0
