"""redis:// backend of the dynamic-embedding PS (csrc/dynemb/redis_io.cpp) against an in-process RESP2 server: the wire format is what
a real Redis speaks (HSET / HMGET / HLEN / AUTH / SELECT), pipelined and chunked. Reference: csrc/dynamic_embedding/details/redis/."""
import socket
import socketserver
import threading

import numpy as np
import pytest
import torch


class _MiniRedis(socketserver.ThreadingTCPServer):
    allow_reuse_address = True
    daemon_threads = True

    def __init__(self, password=None):
        self.dbs = {}
        self.password = password
        self.commands = []
        self.drop_next = 0
        self.lock = threading.Lock()
        super().__init__(("127.0.0.1", 0), _Handler)


class _Handler(socketserver.StreamRequestHandler):
    def _read_cmd(self):
        line = self.rfile.readline()
        if not line:
            return None
        assert line[:1] == b"*", line
        args = []
        for _ in range(int(line[1:])):
            hdr = self.rfile.readline()
            assert hdr[:1] == b"$"
            n = int(hdr[1:])
            args.append(self.rfile.read(n))
            self.rfile.read(2)
        return args

    def handle(self):
        srv = self.server
        db, authed = 0, srv.password is None
        while True:
            try:
                cmd = self._read_cmd()
            except (ConnectionError, AssertionError):
                return
            if cmd is None:
                return
            name = cmd[0].upper()
            with srv.lock:
                srv.commands.append((name, len(cmd)))
                if srv.drop_next > 0:  # simulate a server that closed an idle connection
                    srv.drop_next -= 1
                    self.connection.shutdown(socket.SHUT_RDWR)
                    return
                if name == b"AUTH":
                    authed = cmd[-1].decode() == srv.password
                    self.wfile.write(b"+OK\r\n" if authed else b"-ERR invalid password\r\n")
                elif not authed:
                    self.wfile.write(b"-NOAUTH Authentication required.\r\n")
                elif name == b"SELECT":
                    db = int(cmd[1])
                    self.wfile.write(b"+OK\r\n")
                elif name == b"PING":
                    self.wfile.write(b"+PONG\r\n")
                elif name == b"HSET":
                    h = srv.dbs.setdefault(db, {}).setdefault(cmd[1], {})
                    new = 0
                    for f, v in zip(cmd[2::2], cmd[3::2]):
                        new += f not in h
                        h[f] = v
                    self.wfile.write(b":%d\r\n" % new)
                elif name == b"HMGET":
                    h = srv.dbs.get(db, {}).get(cmd[1], {})
                    out = [b"*%d\r\n" % (len(cmd) - 2)]
                    for f in cmd[2:]:
                        v = h.get(f)
                        out.append(b"$-1\r\n" if v is None else b"$%d\r\n%s\r\n" % (len(v), v))
                    self.wfile.write(b"".join(out))
                elif name == b"HLEN":
                    self.wfile.write(b":%d\r\n" % len(srv.dbs.get(db, {}).get(cmd[1], {})))
                else:
                    self.wfile.write(b"-ERR unknown command\r\n")
            self.wfile.flush()


@pytest.fixture()
def redis_server():
    srv = _MiniRedis(password="s3cret")
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    yield srv
    srv.shutdown()
    srv.server_close()


def test_ps_rows_round_trip_through_redis(redis_server):
    from torchrec_b200.dynamic_embedding.ps import PS

    port = redis_server.server_address[1]
    w = torch.zeros(64, 8)
    m = torch.zeros(64)  # row-wise optimizer state travels with the weight row
    url = f"redis://:s3cret@127.0.0.1:{port}/3?prefix=emb:&chunk=5&pool=2"
    ps = PS("table_a", [w, m], url, init_fn=lambda n: [torch.full((n, 8), -1.0), torch.full((n,), -2.0)])
    w[:20] = torch.arange(20 * 8, dtype=torch.float32).view(20, 8)
    m[:20] = torch.arange(20, dtype=torch.float32) + 0.5
    gids = torch.arange(1000, 1020)
    ps.evict(torch.stack([gids, torch.arange(20)], dim=1))
    ps.wait()
    assert len(ps) == 20
    h = redis_server.dbs[3][b"emb:table_a"]
    assert len(h) == 20 and np.frombuffer(next(iter(h)), dtype=np.int64)[0] in range(1000, 1020)
    assert all(len(v) == 8 * 4 + 4 for v in h.values())
    hsets = [c for c in redis_server.commands if c[0] == b"HSET"]
    assert len(hsets) == 4 and all(n == 2 + 2 * 5 for _, n in hsets), "20 rows in chunks of 5 -> 4 pipelined HSETs"
    # fetch 10 known + 5 unknown ids into other cache slots
    want = torch.cat([gids[5:15], torch.arange(5000, 5005)])
    slots = torch.arange(40, 55)
    ps.fetch(torch.stack([want, slots], dim=1))
    torch.testing.assert_close(w[40:50], w[5:15])
    torch.testing.assert_close(m[40:50], m[5:15])
    assert bool((w[50:55] == -1).all()) and bool((m[50:55] == -2).all())
    # overwrite: the newest blob wins, the table does not grow
    w[5] += 100
    ps.evict(torch.tensor([[1005, 5]]))
    ps.wait()
    assert len(ps) == 20
    ps.fetch(torch.tensor([[1005, 60]]))
    torch.testing.assert_close(w[60], w[5])
    # a connection the server dropped while idle is re-opened transparently (AUTH + SELECT again)
    redis_server.drop_next = 1
    ps.fetch(torch.tensor([[1007, 61]]))
    torch.testing.assert_close(w[61], w[7])


def test_redis_backend_errors_are_loud(redis_server):
    from torchrec_b200.dynamic_embedding.ps import PS

    port = redis_server.server_address[1]
    with pytest.raises(RuntimeError):
        PS("t", [torch.zeros(4, 2)], f"redis://:wrong@127.0.0.1:{port}")
    with pytest.raises(RuntimeError):
        PS("t", [torch.zeros(4, 2)], "redis://127.0.0.1:1?timeout_ms=200")  # nothing listens there
