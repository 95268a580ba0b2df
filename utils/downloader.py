"""Fetch a file over HTTP(S) unless it is already on disk.

Same entry point as the reference's utils/downloader.py (`download_url(url, file_path)`, with the
skip-if-present behaviour examples/mnist/run.py relies on).  The transfer is streamed to a
temporary file and renamed at the end, so an interrupted download never leaves a truncated file
that the next run would mistake for the data set."""
import os
import shutil
import tempfile
from urllib.error import URLError
from urllib.request import urlopen

_CHUNK = 1 << 16


def show_progress(blk_num, blk_sz, tot_sz):
    """progress line; tolerates servers that do not send a length (tot_sz <= 0)"""
    done = blk_num * blk_sz
    if tot_sz and tot_sz > 0:
        print("Progress: %.1f %%" % (100.0 * min(done, tot_sz) / tot_sz), end="\r", flush=True)
    else:
        print("Progress: %d bytes" % done, end="\r", flush=True)


def download_url(url, file_path):
    target_dir = os.path.dirname(file_path)
    if target_dir:
        os.makedirs(target_dir, exist_ok=True)
    if os.path.exists(file_path):
        print("%s already exists." % file_path)
        return file_path
    print("Downloading %s to %s" % (url, file_path))
    fd, tmp_path = tempfile.mkstemp(dir=target_dir or ".", suffix=".part")
    try:
        with os.fdopen(fd, "wb") as sink:
            try:
                with urlopen(url) as source:
                    total = int(source.headers.get("Content-Length") or 0)
                    blocks = 0
                    while True:
                        chunk = source.read(_CHUNK)
                        if not chunk:
                            break
                        sink.write(chunk)
                        blocks += 1
                        show_progress(blocks, _CHUNK, total)
            except URLError:
                raise RuntimeError("Error downloading resource!")
            finally:
                print()
        shutil.move(tmp_path, file_path)
    except KeyboardInterrupt:
        print("Interrupted")
    finally:
        if os.path.exists(tmp_path):
            os.remove(tmp_path)
    return file_path
