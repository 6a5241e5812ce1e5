"""DeepLCD scoring / DetectLoop on the GPU against the numpy restatement (src/deeplcd.cpp:35-39,
src/loopclosing.cpp:124-161).  fp32 database: |score difference| <= 1e-6 (Eigen's summation order
is unspecified); fp16 database (north_star): <= 1e-3, and identical loop decisions on the synthetic database."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(0, 1e-6), (1, 1e-3)])
def test_scores_and_loop_decisions(pkg, oracle, synth, dtype, tol):
    db = synth.lcd_database(0)
    ids = np.arange(len(db)) * 3 + 1                 # keyframe ids need not be dense
    lcd = pkg.DeepLCDScorer(capacity=1024, dtype=dtype, max_queries=8)
    loops = 0
    for k in range(len(db)):
        cur_id = int(ids[k])
        if len(lcd) > 50:                            # LCD.nDatabaseMinSize (src/loopclosing.cpp:66)
            got = lcd.DetectLoop(cur_id, db[k], min_gap=60)
            want = oracle.lcd_detect_loop(ids[:len(lcd)], db[:len(lcd)], cur_id, db[k], min_gap=60)
            assert got[0] == want[0] and got[3] == want[3], (k, got, want)
            assert abs(got[2] - want[2]) <= tol
            if want[0]:
                assert got[1] == want[1]
                loops += 1
        lcd.add(cur_id, db[k])
    assert loops == len(synth.LOOP_PAIRS)
    q = db[[5, 100, 389, 741]]
    s = lcd.score(q)
    want = q.astype(np.float64) @ db.astype(np.float64).T
    assert s.shape == (4, len(db)) and np.abs(s - want).max() <= tol
    lcd.close()


def test_remove_and_recency_break(pkg, oracle, synth):
    db = synth.lcd_database(1, n=120, pairs=[(100, 10)])
    lcd = pkg.DeepLCDScorer(capacity=128, dtype=0)
    lcd.add_batch(np.arange(100), db[:100])
    assert len(lcd) == 100
    found, best, mx, cnt = lcd.DetectLoop(100, db[100])
    assert found and best == 10 and cnt == 1
    lcd.remove(10)                                    # the reference erases a confirmed loop keyframe
    assert len(lcd) == 99
    got = lcd.DetectLoop(100, db[100])
    ids = np.array([i for i in range(100) if i != 10])
    want = oracle.lcd_detect_loop(ids, db[ids], 100, db[100])
    assert got[0] == want[0] == False and got[1] == want[1] and abs(got[2] - want[2]) < 1e-6
    # everything is "recent": nothing is scanned (break at the first row)
    assert lcd.DetectLoop(15, db[100]) == (False, 0, 0.0, 0)
    # quirk Q12: the scan stops at the first recent row even if older rows follow a gap in the ids
    s = lcd.score(db[100:101])
    assert s.shape == (1, 99)
    lcd.close()


@pytest.mark.parametrize("n_db,nq", [(742, 742), (300, 64), (1025, 130), (7, 64)])
def test_batch_scores_tensor_core_path(pkg, oracle, synth, n_db, nq):
    """A batch of >= 64 queries against the fp16 database runs as an fp16 GEMM on tcgen05 (k_lcd_score_umma): every score
    within the fp16 tolerance (1e-3, the bound of the GEMV path) of the fp32 DeepLCD::score, within 5e-4 of the GEMV
    path, and the same loop decisions; ragged tile edges (rows, queries not multiples of 256 / 128)."""
    db = synth.lcd_database(3, n=max(n_db, nq))
    rows, queries = db[:n_db], db[:nq][::-1].copy()
    lcd = pkg.DeepLCDScorer(capacity=n_db + 3, dtype=1, max_queries=nq)
    lcd.add_batch(np.arange(n_db), rows)
    got = lcd.score(queries)                                   # tensor-core path (nq >= 64)
    assert got.shape == (nq, n_db)
    want = (queries.astype(np.float32) @ rows.astype(np.float32).T)
    assert np.abs(got - want).max() <= 1e-3
    gemv = np.stack([lcd.score(queries[i:i + 1])[0] for i in range(0, nq, max(1, nq // 16))])   # GEMV path (one query)
    assert np.abs(got[::max(1, nq // 16)] - gemv).max() <= 5e-4
    if n_db >= 2:   # the best-scoring row (the loop candidate) is the same wherever the top two are further apart than the tolerance
        srt = np.sort(want, 1)
        clear = (srt[:, -1] - srt[:, -2]) > 2e-3
        assert np.array_equal(got.argmax(1)[clear], want.argmax(1)[clear])
