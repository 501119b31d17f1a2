#!/usr/bin/env python
"""oracle/mprun.py -- TEST INFRASTRUCTURE ONLY.  Starts N ranks of a program linked against the multi-process MPI
replacement (oracle/mpi_mp): one Unix socket pair per pair of ranks, handed to the children as inherited descriptors.

usage: python oracle/mprun.py -n N [--timeout S] program [args ...]"""
import os
import socket
import subprocess
import sys


def main():
    argv = sys.argv[1:]
    n, timeout = 1, 600.0
    while argv and argv[0].startswith("-"):
        if argv[0] == "-n":
            n = int(argv[1]); argv = argv[2:]
        elif argv[0] == "--timeout":
            timeout = float(argv[1]); argv = argv[2:]
        else:
            sys.exit(__doc__)
    if not argv:
        sys.exit(__doc__)
    fds = [[-1] * n for _ in range(n)]
    socks = []
    for i in range(n):
        for j in range(i + 1, n):
            a, b = socket.socketpair(socket.AF_UNIX, socket.SOCK_STREAM)
            for s in (a, b):
                s.setsockopt(socket.SOL_SOCKET, socket.SO_SNDBUF, 1 << 22)
                s.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 1 << 22)
            fds[i][j], fds[j][i] = a.fileno(), b.fileno()
            socks += [a, b]
    procs = []
    for r in range(n):
        env = dict(os.environ, SB200_MPI_RANK=str(r), SB200_MPI_SIZE=str(n), SB200_MPI_FDS=",".join(map(str, fds[r])))
        procs.append(subprocess.Popen(argv, env=env, pass_fds=[f for f in fds[r] if f >= 0]))
    for s in socks:
        s.close()
    rc = 0
    try:
        for p in procs:
            rc = p.wait(timeout=timeout) or rc
    except subprocess.TimeoutExpired:
        rc = 124
        for p in procs:
            p.kill()
    sys.exit(rc)


if __name__ == "__main__":
    main()
