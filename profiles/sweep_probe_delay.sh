for d in 0 100 200 300 500 800; do echo "== fwd delay $d"; ASR_LSTM_FWD_DELAY_NS=$d python profiles/prof_lstm_phases.py 2>&1 | grep "fwd ms" | tail -1; done
for d in 0 100 200 300 500 800; do echo "== bwd delay1 $d"; ASR_LSTM_BWD_DELAY1_NS=$d python profiles/prof_lstm_phases.py 2>&1 | grep "bwd ms" | tail -1; done
for d in 100 200 300 500; do echo "== bwd delay2 $d"; ASR_LSTM_BWD_DELAY2_NS=$d python profiles/prof_lstm_phases.py 2>&1 | grep "bwd ms" | tail -1; done
