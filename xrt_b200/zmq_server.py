# -*- coding: utf-8 -*-
"""A CUDA server for xrt's remote-accelerator clients.

xrt's `XRT_CL` turns into a ZMQ client when `targetOpenCL` is 'host:port' (xrt/backends/raycing/myopencl.py:165-203):
`run_parallel` then pickles its keyword arguments (protocol 2) into a REQ socket and unpickles the reply
(:421-436).  The reference's server (tests/raycing/RemoteOpenCLCalculation/zmq_server.py:115-199) answers from pyopencl
contexts; this one answers from `XRT_CUDA.run_parallel`, same wire protocol, same kernel names, the 'PONG' reply to the
client's ping included -- an unmodified xrt client with targetOpenCL='127.0.0.1:15559' computes on the B200.

Two things differ from the reference server, on purpose:
  * it binds to the loopback interface unless told otherwise (the reference binds tcp://*: anyone who can reach the
    port can make it unpickle arbitrary objects);
  * requests are decoded by a RESTRICTED unpickler that only rebuilds what a run_parallel call consists of -- dicts,
    lists, tuples, strings, numbers, None and numpy arrays / scalars / dtypes; anything else in the stream is an
    error reply, never an import.

    python -m xrt_b200.zmq_server [--bind tcp://127.0.0.1:15559] [--device 0]
"""
import argparse
import io
import pickle

import numpy as np

PING_KERNEL = '__xrt_ping__'  # myopencl._ZMQ_PING_KERNEL (myopencl.py:36)
DEFAULT_BIND = 'tcp://127.0.0.1:15559'  # zmq_server.py:131 uses port 15559 in standalone mode

_ALLOWED = {
    ('numpy.core.multiarray', '_reconstruct'), ('numpy._core.multiarray', '_reconstruct'),
    ('numpy.core.multiarray', 'scalar'), ('numpy._core.multiarray', 'scalar'),
    ('numpy', 'ndarray'), ('numpy', 'dtype'),
    ('numpy.core.numeric', '_frombuffer'), ('numpy._core.numeric', '_frombuffer'),
    ('builtins', 'complex'), ('__builtin__', 'complex'), ('builtins', 'bytearray'), ('_codecs', 'encode'),
}


class _NumpyOnlyUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError('refused: {0}.{1} is not part of a run_parallel message'.format(module, name))


def loads(data):
    return _NumpyOnlyUnpickler(io.BytesIO(data), encoding='latin1').load()


def dumps(obj):
    return pickle.dumps(obj, 2)  # the client reads protocol 2 (myopencl.py send/recv_zipped_pickle)


def handle(backend, data):
    """One request -> one reply (bytes -> bytes).  Errors travel as ("ERROR", text) like the reference's."""
    try:
        message = loads(data)
    except Exception as e:  # malformed or refused stream
        return dumps(('ERROR', 'cannot decode the request: {0}'.format(e)))
    if not isinstance(message, dict) or 'kernelName' not in message:
        return dumps(('ERROR', 'not a kernel'))
    if message.get('kernelName') == PING_KERNEL:
        return dumps(('PONG',))
    allowed = ('kernelName', 'scalarArgs', 'slicedROArgs', 'nonSlicedROArgs', 'slicedRWArgs', 'nonSlicedRWArgs',
               'dimension', 'complexity')
    kwargs = {k: message[k] for k in allowed if k in message}
    for arg in kwargs.get('scalarArgs') or []:
        if isinstance(arg, np.float32):
            return dumps(('ERROR', 'float32 requests are not served: xrt_b200 computes in float64 only'))
    try:
        reply = backend.run_parallel(**kwargs)
    except NotImplementedError as e:
        return dumps(('ERROR', 'unknown kernel: {0}'.format(e)))
    except Exception as e:
        return dumps(('ERROR', 'error while calculating a kernel: {0}'.format(e)))
    return dumps(tuple(reply))


def serve(bind=DEFAULT_BIND, device=None, backend=None, max_requests=None):
    import zmq
    if backend is None:
        from .cuda_backend import XRT_CUDA
        backend = XRT_CUDA(targetOpenCL='CUDA' if device is None else 'CUDA:%d' % device)
    context = zmq.Context.instance()
    socket = context.socket(zmq.REP)
    socket.bind(bind)
    served = 0
    try:
        while max_requests is None or served < max_requests:
            socket.send(handle(backend, socket.recv()))
            served += 1
    finally:
        socket.close(linger=0)
    return served


def main():
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--bind', default=DEFAULT_BIND, help='ZMQ endpoint to bind (loopback by default)')
    ap.add_argument('--device', type=int, default=None)
    args = ap.parse_args()
    print('xrt_b200 CUDA server on', args.bind)
    serve(args.bind, args.device)


if __name__ == '__main__':
    main()
