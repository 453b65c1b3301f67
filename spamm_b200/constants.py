"""Physical constants as the reference holds them: astropy 3.x => CODATA-2014,
converted the way astropy does (value * unit scale).  Call sites:
BalmerContinuumCombined.py:10,21-26; FeComponent.py:270; HostGalaxyComponent.py:275."""
import numpy as np

C_SI = 299792458.0                 # m / s
H_SI = 6.626070040e-34             # J s
KB_SI = 1.38064852e-23             # J / K
RYD_SI = 10973731.568508           # 1 / m

C_KMS = C_SI / 1000.0              # c.to("km/s").value
C_CGS = C_SI * 100.0               # c.cgs.value  [cm/s]
C_AA = C_SI * 1e10                 # Angstrom / s
H_CGS = H_SI * 1e7                 # erg s
KB_CGS = KB_SI * 1e7               # erg / K
RYD_AA = RYD_SI * 1e-10            # Ryd.to("1/Angstrom").value
E0 = 2.179e-11                     # BalmerContinuumCombined.py:25
BALMER_EDGE = 3646                 # BalmerContinuumCombined.py:26 (Angstrom)

FWHM_TO_SIGMA_DENOM = C_KMS * 2. * np.sqrt(2. * np.log(2.))   # Fe:286 / HG:293
SQRT_2PI = float(np.sqrt(2 * np.pi))                         # Fe:293
