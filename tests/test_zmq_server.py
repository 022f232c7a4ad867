"""The CUDA stand-in of xrt's remote OpenCL server (tests/raycing/RemoteOpenCLCalculation/zmq_server.py): wire protocol
with a client that does exactly what XRT_CL does in ZMQ mode (myopencl.py:165-203, :421-436), restricted decoding."""
import pickle
import threading

import numpy as np
import pytest

zmq = pytest.importorskip('zmq')


class _FakeBackend(object):
    def run_parallel(self, kernelName='', scalarArgs=None, slicedROArgs=None, nonSlicedROArgs=None, slicedRWArgs=None,
                     nonSlicedRWArgs=None, dimension=0, complexity=0, signal=None):
        if kernelName != 'double_it':
            raise NotImplementedError(kernelName)
        slicedRWArgs[0][:] = 2 * slicedROArgs[0] + scalarArgs[0]
        return tuple(slicedRWArgs)


def _client(endpoint, message):
    ctx = zmq.Context.instance()
    s = ctx.socket(zmq.REQ)
    s.setsockopt(zmq.LINGER, 0)
    s.setsockopt(zmq.RCVTIMEO, 20000)
    s.connect(endpoint)
    s.send(pickle.dumps(message, 2) if not isinstance(message, bytes) else message)
    ret = pickle.loads(s.recv())
    s.close()
    return ret


def test_zmq_roundtrip_ping_kernel_and_refusals():
    from xrt_b200 import zmq_server
    endpoint = 'tcp://127.0.0.1:15587'
    th = threading.Thread(target=zmq_server.serve, kwargs=dict(bind=endpoint, backend=_FakeBackend(), max_requests=5))
    th.daemon = True
    th.start()
    assert _client(endpoint, {'kernelName': zmq_server.PING_KERNEL}) == ('PONG',)
    x = np.arange(7, dtype=np.float64)
    msg = dict(kernelName='double_it', scalarArgs=[np.float64(3.)], slicedROArgs=[x], nonSlicedROArgs=None,
               slicedRWArgs=[np.zeros(7)], nonSlicedRWArgs=None, dimension=7, complexity=0)
    ret = _client(endpoint, msg)
    assert isinstance(ret, tuple) and np.array_equal(ret[0], 2 * x + 3)
    assert _client(endpoint, dict(msg, kernelName='nope'))[0] == 'ERROR'
    assert _client(endpoint, dict(msg, scalarArgs=[np.float32(1.)]))[0] == 'ERROR'  # fp64 only

    class Evil(object):
        def __reduce__(self):
            import os
            return (os.system, ('echo pwned > /tmp/xrt_b200_pwned',))
    ret = _client(endpoint, pickle.dumps({'kernelName': 'double_it', 'scalarArgs': [Evil()]}, 2))
    assert ret[0] == 'ERROR' and 'refused' in ret[1]
    th.join(timeout=10)
    import os
    assert not os.path.exists('/tmp/xrt_b200_pwned')
    assert zmq_server.DEFAULT_BIND.startswith('tcp://127.0.0.1')
