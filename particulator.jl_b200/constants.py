"""Physical constants, values copied bit-for-bit from the reference's `Constants` module
(src/constants.jl, CODATA-2014 via scipy).  Derived constants use the same expression order
(src/constants.jl:157-164) so that host-built tables match a Julia host to the last bit."""
import math

c = 299792458.0                       # constants.jl:37
e = 1.6021766208e-19                  # constants.jl:50
eV = 1.6021766208e-19                 # constants.jl:51
electron_mass = 9.10938356e-31        # constants.jl:52
elementary_charge = 1.6021766208e-19  # constants.jl:54
epsilon_0 = 8.854187817620389e-12     # constants.jl:55
fine_structure = 0.0072973525664      # constants.jl:60
hbar = 1.0545718001391127e-34         # constants.jl:77
N_A = 6.022140857e+23                 # constants.jl:14
kilo = 1000.0                         # constants.jl:86
centi = 0.01                          # constants.jl:42
nair = 2.6867811e+25                  # constants.jl:149
Td = 1e-21                            # constants.jl:150
pi = math.pi

r_e = elementary_charge ** 2 / (electron_mass * c ** 2) / (4 * math.pi * epsilon_0)  # :157
a_0 = hbar / (electron_mass * c * fine_structure)                                     # :160
electron_mc2 = electron_mass * c ** 2                                                 # :163
electron_mc = electron_mass * c                                                       # :164
