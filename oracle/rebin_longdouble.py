"""TEST INFRASTRUCTURE ONLY.  `rebin_spec` (utils/rebin_spec.py:3-17 of the reference, i.e. pysynphot's
`Observation(spec, filt, binset=wavnew, force='taper').binflux`) restated literally, bin by bin, in extended
precision -- the independent check of the product's vectorised / sparse-matrix form
(spamm_b200/utils/rebin_spec.py).  PARITY UNPINNED: pysynphot is absent from this image and the reference
holds no vector for this branch; the algorithm is pysynphot's published `initbinflux`
(merge the native wavelengths with the bin edges and centres, sample the spectrum at the merged nodes,
trapezoid per bin, divide by the bin's summed widths)."""
import numpy as np

LD = np.longdouble


def rebin_spec_longdouble(wave, specin, wavnew):
    wave = np.asarray(wave, dtype=LD)
    f = np.asarray(specin, dtype=LD)
    c = np.asarray(wavnew, dtype=LD)
    edges = np.empty(len(c) + 1, dtype=LD)                      # binning.calculate_bin_edges
    edges[1:-1] = (c[1:] + c[:-1]) / LD(2)
    edges[0] = LD(2) * c[0] - edges[1]
    edges[-1] = LD(2) * c[-1] - edges[-2]
    spwave = np.unique(np.concatenate([wave, edges, c]))       # MergeWaveSets(wave, endpoints), then binwave

    def sample(x):                                              # spectrum x unit throughput, zero outside
        if x < wave[0] or x > wave[-1]:
            return LD(0)
        j = int(np.searchsorted(wave, x, side="right")) - 1
        j = min(max(j, 0), len(wave) - 2)
        t = (x - wave[j]) / (wave[j + 1] - wave[j])
        return f[j] * (LD(1) - t) + f[j + 1] * t

    flux = np.array([sample(x) for x in spwave], dtype=LD)
    avflux = (flux[1:] + flux[:-1]) / LD(2)
    deltaw = spwave[1:] - spwave[:-1]
    indices = np.searchsorted(spwave, edges)
    out = np.empty(len(c), dtype=LD)
    for i in range(len(c)):
        first, last = indices[i], indices[i + 1]
        out[i] = (avflux[first:last] * deltaw[first:last]).sum() / deltaw[first:last].sum()
    return out
