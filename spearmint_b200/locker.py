"""Advisory lock files compatible with the reference's Locker (Locker.py:32-69): a lock on ``f`` is the
symlink ``f.lock -> /dev/null``; creation is atomic, so a reference process and this plugin exclude each other
on the chooser state pickle.  Re-entrant per Locker instance, spin-wait at 10 ms like the reference."""
import os
import sys
import time


class Locker(object):
    def __init__(self):
        self.locks = {}

    def __del__(self):
        for filename in list(self.locks.keys()):
            self.locks[filename] = 1
            self.unlock(filename)

    def lock(self, filename):
        if filename in self.locks:
            self.locks[filename] += 1
            return True
        try:
            os.symlink("/dev/null", "%s.lock" % filename)
        except OSError:
            return False
        self.locks[filename] = 1
        return True

    def unlock(self, filename):
        if filename not in self.locks:
            return True
        if self.locks[filename] > 1:
            self.locks[filename] -= 1
            return True
        ok = True
        try:
            os.rename("%s.lock" % filename, "%s.lock.delete" % filename)
            os.remove("%s.lock.delete" % filename)
        except OSError:
            ok = False
            sys.stderr.write("Could not unlock file: %s.\n" % filename)
        del self.locks[filename]
        return ok

    def lock_wait(self, filename):
        while not self.lock(filename):
            time.sleep(0.01)


def log(*args):
    """helpers.log of the reference (helpers.py:10-14): message to stderr."""
    for v in args:
        sys.stderr.write(str(v))
    sys.stderr.write("\n")
