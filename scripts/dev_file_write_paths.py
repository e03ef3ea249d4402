import os, time, threading, mmap
import numpy as np
N = 64
buf = np.frombuffer(os.urandom(10 << 20), dtype=np.uint8)
def busy(stop):
    a = np.empty(64 << 20, np.uint8); b = np.empty(64 << 20, np.uint8)
    while not stop[0]:
        b[:] = a
for load in (0, 12):
    stop = [False]
    th = [threading.Thread(target=busy, args=(stop,)) for _ in range(load)]
    [t.start() for t in th]
    for path in ("/dev/shm/y", "/tmp/y"):
        fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
        t0 = time.time()
        for i in range(N):
            os.pwrite(fd, buf, i * buf.size)
        dt = time.time() - t0
        os.close(fd); os.remove(path)
        print("load %2d %s pwrite 1 thread: %.2f GB/s" % (load, path, N * buf.size / dt / 1e9), flush=True)
        for nt in (1, 4, 8):
            fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
            t0 = time.time()
            for i in range(N):
                off = i * buf.size
                os.ftruncate(fd, off + buf.size)
                mm = mmap.mmap(fd, buf.size, mmap.MAP_SHARED, mmap.PROT_WRITE | mmap.PROT_READ, offset=off)
                dst = np.frombuffer(mm, dtype=np.uint8)
                step = buf.size // nt
                def cp(k):
                    dst[k * step:(k + 1) * step] = buf[k * step:(k + 1) * step]
                ts = [threading.Thread(target=cp, args=(k,)) for k in range(nt)]
                [t.start() for t in ts]; [t.join() for t in ts]
                del dst
                mm.close()
            dt = time.time() - t0
            os.close(fd); os.remove(path)
            print("load %2d %s ftruncate+mmap+memcpy x%d: %.2f GB/s" % (load, path, nt, N * buf.size / dt / 1e9), flush=True)
    stop[0] = True
    [t.join() for t in th]
