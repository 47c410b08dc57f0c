"""Fortran namelist writer with the interface of the reference's namelist.py (write_namelist)."""


def write_namelist(fname, names, args, vals, nargs):
    with open(fname, 'w') as f:
        indx = 0
        for i in range(len(names)):
            f.write('&' + names[i] + '\n')
            for _ in range(nargs[i]):
                f.write(' ' + args[indx] + '=' + str(vals[indx]) + ',\n')
                indx += 1
            f.write('/\n')
