"""Flux-loop helpers with the reference's file format (ThinCurr/sensor.py:95-107, floops.loc)."""
import numpy


def circular_flux_loop(R, Z, name, scale=1.0, npts=180):
    theta = numpy.linspace(0.0, 2.0 * numpy.pi, npts)
    pts = numpy.stack([R * numpy.cos(theta), R * numpy.sin(theta), Z * numpy.ones(npts)], 1)
    return dict(name=name, scale=scale, pts=pts)


def save_sensors(sensors, filename='floops.loc'):
    with open(filename, 'w+') as fid:
        fid.write('{0}\n'.format(len(sensors)))
        for s in sensors:
            fid.write('\n{0} {1:.6E} {2}\n'.format(s['pts'].shape[0], s['scale'], s['name']))
            for p in s['pts']:
                fid.write('{0:.6E} {1:.6E} {2:.6E}\n'.format(*p))
