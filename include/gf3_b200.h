/* gf3_b200.h -- C ABI of the B200-native GF3 OFDM physical layer (libgf3b200.so).
 *
 * The reference (adamg-97/GF3-audio-modem) is one Python module, OFDM.py, with no FFI of its
 * own; its boundary is the numpy-in / numpy-out method surface of the classes CamG /
 * transmitter / receiver.  This header is the C-ABI a binding for that surface loads (the
 * ctypes stub is gf3-audio-modem_b200/gf3b200/_lib.py; INTEGRATION.md shows the reference-side
 * patch).  Each entry point names the reference lines (file:line in the reference repo) it
 * replaces.
 *
 * Conventions
 *  - every array pointer is a DEVICE pointer owned by the caller (e.g. torch tensor storage);
 *    the library allocates nothing except inside an explicit gf3_plan handle (small constant
 *    tables: FFT twiddles, chirp spectrum);
 *  - every call is asynchronous on the cudaStream_t passed as `stream` (NULL = default stream);
 *  - return value 0 = OK, < 0 = error; gf3_last_error() returns a thread-local message;
 *  - complex arrays are interleaved float (re, im) = numpy complex64;
 *  - "packed bits" are MSB-first bytes exactly as np.packbits produces (OFDM.py:761,769);
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GF3_B200_H
#define GF3_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GF3_ABI_VERSION 1

/* exported from libgf3b200.so (everything else in the library has hidden visibility) */
#if defined(__GNUC__)
#define GF3_API __attribute__((visibility("default")))
#else
#define GF3_API
#endif

/* sample formats of received audio: float32, or PCM as recorded (Final System Test.ipynb:85-86 reads an
 * 8-bit wav and converts with r/1.0): plain value conversion, no scaling, no DC removal */
enum { GF3_SAMPLE_U8 = 0, GF3_SAMPLE_I16 = 1, GF3_SAMPLE_F32 = 2 };

enum {
    GF3_OK = 0,
    GF3_ERR_INVALID = -1,   /* bad argument / unsupported parameter combination */
    GF3_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
    GF3_ERR_NODEVICE = -3   /* no CUDA device: the product has no CPU path */
};

/* Parameter contract of CamG.__init__ (OFDM.py:18-101), generalised to any power-of-two N. */
typedef struct gf3_params {
    int32_t N;            /* OFDM.py:27   ofdm_symbol_size (power of two, 64..4096)        */
    int32_t cp;           /* OFDM.py:42   cyclic prefix length (even)                      */
    int32_t lo;           /* OFDM.py:43   lowest data bin, inclusive (>= 1)                */
    int32_t hi;           /* OFDM.py:44,47 highest data bin, EXCLUSIVE (<= N/2)            */
    int32_t n_pilots;     /* OFDM.py:51   known OFDM symbols before and after the data     */
    int32_t packet_len;   /* OFDM.py:50   data symbols per packet                          */
    int32_t fit_lo;       /* OFDM.py:462  phase-slope fit window start (reference: 500)    */
    int32_t fit_hi;       /* OFDM.py:462  phase-slope fit window end, exclusive (1000)     */
    int32_t chirp_len;    /* OFDM.py:64   5*(N+cp) in the reference                        */
    float fs;             /* OFDM.py:24   48000                                            */
    float f0;             /* OFDM.py:62   chirp start frequency                            */
    float f1;             /* OFDM.py:63   chirp end frequency                              */
    float thresh;         /* OFDM.py:361  0.4 (fraction of the global matched-filter max)  */
    float tx_gain;        /* OFDM.py:256  2.0                                              */
    float chirp_gain;     /* OFDM.py:109  0.2 (= 1/5)                                      */
} gf3_params;

typedef struct gf3_plan gf3_plan;   /* opaque: device-resident constant tables for one params set */

/* ---- library / plan ------------------------------------------------------------------ */
GF3_API int gf3_abi_version(void);
GF3_API const char* gf3_last_error(void);
/* Number of CUDA devices visible; 0 means every compute call will return GF3_ERR_NODEVICE. */
GF3_API int gf3_device_count(void);
/* Fill derived defaults (fit window 500..1000, chirp_len 5*(N+cp), fs 48000, f0 0, f1 8000,
 * thresh 0.4, tx_gain 2, chirp_gain 0.2) -- the constants hard-coded in OFDM.py:24,62-64,109,
 * 256,361,462.  Only N, cp, lo, hi, n_pilots, packet_len are read from the arguments. */
GF3_API int gf3_params_default(gf3_params* p, int N, int cp, int lo, int hi, int n_pilots, int packet_len);
/* Builds the twiddle tables on the current device.  Replaces CamG.__init__ (OFDM.py:18-101). */
GF3_API int gf3_plan_create(const gf3_params* p, gf3_plan** out);
GF3_API int gf3_plan_destroy(gf3_plan* plan);
GF3_API int gf3_plan_params(const gf3_plan* plan, gf3_params* out);
/* Kernels launched by this library since load (all entry points; for bench.py's gpu_launches). */
GF3_API int64_t gf3_launch_count(void);

/* ---- receive chain ----------------------------------------- */
/* Packet p's samples start at samples[pkt_offset[p]] (pkt_offset == NULL: packets are
 * contiguous, offset p*(2P+L)*(N+cp)).  This one addressing mode covers both the reference's
 * get_symbols slicing (OFDM.py:391-403: offset = sync index) and pre-sliced symbol arrays.
 * Any sample alignment is accepted; 8-byte aligned packets take the vectorised load path.   */

/* Channel estimate from the known symbols: remove_cp + fft + get_data (pilot part) + the first
 * half of equalise (OFDM.py:407-418, 593, 429-462).
 *   known[K]   complex64, the known QPSK symbols on bins 1..K (OFDM.py:429)
 *   Hs, He     complex64 [n_packets, K]   = mean_P(FFT(pilots))/known        (OFDM.py:443-451)
 *   slope      float64   [n_packets]      = LS slope of unwrap(angle(He))-unwrap(angle(Hs))
 *                                           over bins [fit_lo, min(fit_hi,K))  (OFDM.py:454-462) */
GF3_API int gf3_rx_estimate(const gf3_plan* plan, const float* samples, const int64_t* pkt_offset,
                    int64_t n_packets, const float* known, float* Hs, float* He, double* slope,
                    void* stream);

/* Data symbols: remove_cp + fft + get_data + equalise second half + carrier select + demap + PS
 * + decode("XOR") fused; every data sample is read from HBM once (OFDM.py:407-418, 593,
 * 466-478, 603, 484-505, 541-544).
 *   xor2[Nd]     uint8 or NULL: per data carrier (b0<<1)|b1 of known_sequence[:2Nd] (OFDM.py:542)
 *   bits_packed  uint8 [n_packets, bits_stride]: packet p's L*Nd*2 bits, MSB first, symbol-major,
 *                carrier, b0 then b1 (OFDM.py:500,505); bits_stride % 4 == 0 and
 *                bits_stride*8 >= L*Nd*2 rounded up to 32; pad bits are written as 0
 *   eq           complex64 [n_packets, L, K] or NULL: the equalised constellation of all K bins
 *                (OFDM.py:478,480), for parity checks; off for throughput                   */
GF3_API int gf3_rx_demod(const gf3_plan* plan, const float* samples, const int64_t* pkt_offset,
                 int64_t n_packets, const float* Hs, const float* He, const double* slope,
                 const uint8_t* xor2, uint8_t* bits_packed, int64_t bits_stride, float* eq,
                 void* stream);

/* The whole receive chain of a batch of packets in ONE launch: gf3_rx_estimate + gf3_rx_demod fused
 * (every persistent CTA estimates the channel of a packet when it first touches it), so each received
 * sample is read from HBM once and the estimate's memory-bound phase overlaps the data symbols'
 * compute of the co-resident CTAs (OFDM.py:591-609 = receiver.receive after synchronisation).
 * Hs, He, slope are OUTPUTS; the other arguments are those of gf3_rx_demod.                       */
GF3_API int gf3_rx_receive(const gf3_plan* plan, const float* samples, const int64_t* pkt_offset,
                   int64_t n_packets, const float* known, float* Hs, float* He, double* slope,
                   const uint8_t* xor2, uint8_t* bits_packed, int64_t bits_stride, float* eq,
                   void* stream);
/* 1 if gf3_rx_receive runs as one launch for this plan; 0 if it falls back to the two launches
 * gf3_rx_estimate + gf3_rx_demod (N = 4096, where the in-kernel estimate would be slower, or a fit
 * window too wide for the kernel's scratch).  The results are the same either way.               */
GF3_API int gf3_rx_receive_is_fused(const gf3_plan* plan);

/* gf3_rx_receive on samples in their recorded format (sample_format: GF3_SAMPLE_U8 / _I16 / _F32; pkt_offset and
 * the contiguous packet stride count SAMPLES).  The symbols enter the kernels through a shared-memory staging
 * buffer filled by bulk asynchronous copies (cp.async.bulk, completion on an mbarrier) one batch ahead of the
 * FFT and are converted in registers, so PCM recordings (Final System Test.ipynb:85-86: an 8-bit wav) cost 1 or
 * 2 bytes of HBM traffic per sample and no float copy of the recording exists; packets may start at any sample.
 * uint8 samples have their offset of 128 removed (it only reaches FFT bin 0, which the chain never reads). */
GF3_API int gf3_rx_receive_pcm(const gf3_plan* plan, const void* samples, int32_t sample_format, const int64_t* pkt_offset,
                       int64_t n_packets, const float* known, float* Hs, float* He, double* slope,
                       const uint8_t* xor2, uint8_t* bits_packed, int64_t bits_stride, float* eq, void* stream);

/* Known-channel receive (old API, Weekend Challenge.ipynb:162-226): Y/H on bins 1..K, demap.
 * Uses the plan's geometry with its n_pilots leading/trailing symbols skipped (create the plan
 * with n_pilots = 0, packet_len = symbols per block).  Hinv[K] = 1/H on bins 1..K.          */
GF3_API int gf3_rx_known_channel(const gf3_plan* plan, const float* samples, const int64_t* pkt_offset,
                         int64_t n_packets, const float* Hinv, const uint8_t* xor2,
                         uint8_t* bits_packed, int64_t bits_stride, float* eq, void* stream);

/* Spectrum of arbitrary symbols: remove_cp + fft + bins 1..K (OFDM.py:407-408, 593, 414-416).
 *   sym_offset[n_symbols] sample index of each symbol's first CP sample (NULL: contiguous)
 *   out complex64 [n_symbols, K]                                                            */
GF3_API int gf3_rx_spectrum(const gf3_plan* plan, const float* samples, const int64_t* sym_offset,
                    int64_t n_symbols, float* out, void* stream);

/* ---- chirp synchronisation ------------------------------------- */
/* sync_chirp (OFDM.py:106-109): float32 [chirp_len]. */
GF3_API int gf3_sync_chirp(const gf3_plan* plan, float* out, void* stream);
/* Matched filter, the convolution of chirp_method (OFDM.py:357-358):
 *   r [n_streams, T] float32 (row stride r_stride samples)
 *   P [n_streams, T + chirp_len - 1] float32 (row stride p_stride), full linear convolution
 *     with the time-reversed chirp, computed by overlap-save FFT blocks;
 *   pmax [n_streams] float32: the signed global maximum of each row (np.amax, OFDM.py:359).
 *   work: device scratch of gf3_xcorr_work_bytes() bytes.                                   */
GF3_API size_t gf3_xcorr_work_bytes(const gf3_plan* plan, int64_t n_streams, int64_t T);
GF3_API int gf3_xcorr(const gf3_plan* plan, const float* r, int64_t r_stride, int64_t n_streams,
              int64_t T, float* P, int64_t p_stride, float* pmax, void* work, void* stream);
/* Peak picking of chirp_method + get_symbols (OFDM.py:359-372, 393-395): normalise by pmax,
 * candidate mask (D[i]*D[i+1] <= 0) & (P[i+1] > thresh), ascending hold-off of chirp_len,
 * including the end-of-signal wipe-out quirk (OFDM.py:366-370).
 *   peaks [n_streams, max_peaks] int64: indices into the reference's `zeros` array
 *   count [n_streams] int32: detections found (may exceed max_peaks; only max_peaks stored)
 *   work: device scratch of gf3_peak_pick_work_bytes() bytes (one candidate bit per position) */
GF3_API size_t gf3_peak_pick_work_bytes(const gf3_plan* plan, int64_t n_streams, int64_t T);
GF3_API int gf3_peak_pick(const gf3_plan* plan, const float* P, int64_t p_stride, int64_t n_streams,
                  int64_t T, const float* pmax, int64_t* peaks, int32_t max_peaks, int32_t* count,
                  void* work, void* stream);

/* chirp_method (OFDM.py:356-372) for a batch of streams in one call: gf3_xcorr + gf3_peak_pick, reading the
 * samples in their native format (sample_format: GF3_SAMPLE_*; r_stride in samples).  When the chirp spans at
 * most four 2048-sample partitions the matched filter runs as ONE fused kernel (forward FFT, partition
 * multiply-accumulate and inverse FFT with the spectra kept on chip) that also records the maximum of every
 * 2048-sample block of P, and the detection walk then reads P only inside blocks that can reach the threshold.
 * P, pmax, peaks, count as in gf3_xcorr / gf3_peak_pick; work: gf3_sync_work_bytes() bytes of device scratch. */
GF3_API size_t gf3_sync_work_bytes(const gf3_plan* plan, int64_t n_streams, int64_t T);
GF3_API int gf3_sync_streams(const gf3_plan* plan, const void* r, int32_t sample_format, int64_t r_stride, int64_t n_streams,
                     int64_t T, float* P, int64_t p_stride, float* pmax, int64_t* peaks, int32_t max_peaks,
                     int32_t* count, void* work, void* stream);

/* The same detections when the caller does not need P itself (chirp_method returns only `zeros`): P_scratch has the
 * size and stride of P but only the blocks the detection rule can read are computed.  A candidate needs
 * P[i+1] > thresh * max(P) (OFDM.py:361) and every sample of a 2048-sample block is bounded by the l1 norm of the
 * block's spectrum, so a block whose bound stays below thresh * (the largest sample seen so far in its stream) is not
 * transformed back to the time domain.  peaks / count / pmax are identical to gf3_sync_streams' (tested); how much
 * work is skipped depends on the recording (88 % of the inverse transforms on the C3 framing at 20 dB, 83 % at 8 dB, none below ~5 dB). */
GF3_API int gf3_sync_detect(const gf3_plan* plan, const void* r, int32_t sample_format, int64_t r_stride, int64_t n_streams,
                    int64_t T, float* P_scratch, int64_t p_stride, float* pmax, int64_t* peaks, int32_t max_peaks,
                    int32_t* count, void* work, void* stream);

/* receiver.schmidlcox_method (OFDM.py:376-387; unused by receive() since the chirp became the standard, still a public
 * method): the timing metric P[d+1] = P[d] + r[d+L] r[d+2L] - r[d] r[d+L], L = N/2, over the first `search` samples as a
 * double-precision prefix sum; index[s] = first argmax |P| (the reference returns index + N - 1), value[s] = |P| there
 * (or NULL).  Needs T >= search - 1 + 2L samples per stream (the reference raises IndexError otherwise). */
GF3_API int gf3_schmidlcox(const gf3_plan* plan, const void* r, int32_t sample_format, int64_t r_stride, int64_t n_streams,
                   int64_t T, int64_t search, int64_t* index, double* value, void* stream);

/* get_symbols' index bookkeeping (OFDM.py:393-397) for a batch of streams, on the device:
 * zero_indicies = where(zeros) + 2 with the last detection (the terminating chirp) dropped, turned
 * into packet offsets for gf3_rx_receive: pkt_offset[s*pk_expected + j] = s*r_stride + peaks[s][j] + 2.
 *   ok [n_streams] uint8 or NULL: 1 when stream s holds exactly pk_expected packets (count ==
 *   pk_expected + 1) and the last one ends inside its T samples -- where the reference's vstack /
 *   reshape would succeed (OFDM.py:400-403); otherwise 0 and the offsets are clamped into the stream. */
GF3_API int gf3_peaks_to_offsets(const gf3_plan* plan, const int64_t* peaks, const int32_t* count, int64_t n_streams,
                         int32_t max_peaks, int64_t r_stride, int64_t T, int32_t pk_expected,
                         int64_t* pkt_offset, uint8_t* ok, void* stream);

/* ---- transmit chain ---------------------------------------------- */
/* map + build_OFDM_symbol + ifft + add_cp + send_to_stream fused (OFDM.py:191-226, 244-259,
 * 322-323).  One launch writes whole packets [chirp | g*(P x known) | g*(L x data) | g*(P x
 * known)] and, per stream, one trailing chirp.
 *   bits_packed uint8 [n_streams, pk_per_stream, bits_stride]  encoded bits (after
 *               transmitter.encode), MSB first, L*Nd*2 bits per packet
 *   filler  complex64 [n_streams, K-Nd]  QPSK on the unused bins (random_qpsk, OFDM.py:201-203)
 *   known   complex64 [K]
 *   out     float32 [n_streams, out_stride], out_stride >= pk_per_stream*(chirp_len +
 *           (2P+L)(N+cp)) + chirp_len                                                       */
GF3_API int gf3_tx_modulate(const gf3_plan* plan, const uint8_t* bits_packed, int64_t bits_stride,
                    const float* filler, const float* known, int64_t n_streams,
                    int64_t pk_per_stream, float* out, int64_t out_stride, void* stream);

/* gf3_tx_modulate with transmitter.encode("XOR") fused (OFDM.py:163-166): bits_packed holds the UN-encoded bits and every data
 * carrier's bit pair is XORed with xor2[carrier] = (b0 << 1) | b1 of known_sequence[:2 Nd] while the symbol is built (the mirror of
 * the receive kernels' fused decode).  The reference pads AFTER encoding; a caller that pads before passes padding already XORed
 * with the same table (x ^ k ^ k = x), as the drop-in's transmit() does. */
GF3_API int gf3_tx_encode_modulate(const gf3_plan* plan, const uint8_t* bits_packed, int64_t bits_stride, const uint8_t* xor2,
                           const float* filler, const float* known, int64_t n_streams,
                           int64_t pk_per_stream, float* out, int64_t out_stride, void* stream);
/* Hermitian spectra -> real symbols: np.fft.ifft(X).real of n_symbols symbols given as bins 1..K (X[0] = X[N/2] = 0,
 * X[N-k] = conj X[k]) with the plan's cyclic prefix prepended and NO transmit gain (OFDM.py:322-323; the old API's module-level
 * IFFT, Initial OFDM Test.ipynb cell 13).  spectrum complex64 [n_symbols, K] -> out float32 [n_symbols, N + cp]. */
GF3_API int gf3_tx_ifft(const gf3_plan* plan, const float* spectrum, int64_t n_symbols, float* out, void* stream);
/* Y / H with H broadcast over the rows: the old API's module-level equalise(Y, H) (Weekend Challenge.ipynb:225).
 * Y, out complex64 [n_rows, m]; H complex64 [m]. */
GF3_API int gf3_cdiv(const float* Y, const float* H, int64_t n_rows, int32_t m, float* out, void* stream);

/* ---- stage-level entry points (SURVEY 8a rows 5, 9, 11 as separate public methods) ------ */
/* receiver.equalise, first half (OFDM.py:429-462) on spectra the caller already holds:
 *   start, end  complex64 [n_packets, P, K]  bins 1..K of the leading / trailing known symbols
 *   -> Hs, He complex64 [n_packets, K], slope float64 [n_packets] exactly as gf3_rx_estimate   */
GF3_API int gf3_eq_estimate(const gf3_plan* plan, const float* start, const float* end, int64_t n_packets,
                    const float* known, float* Hs, float* He, double* slope, void* stream);
/* receiver.equalise, second half (OFDM.py:466-478):
 *   data complex64 [n_packets, L, K] -> eq = data / Hest, Hest = (|Hs| + (|He|-|Hs|) w) *
 *   exp(j (angle(Hs) + slope n w)), w = (l + P/2)/(L + P); Hest [n_packets, L, K] or NULL        */
GF3_API int gf3_eq_apply(const gf3_plan* plan, const float* data, int64_t n_packets, const float* Hs,
                 const float* He, const double* slope, float* eq, float* Hest, void* stream);
/* receiver.demap (OFDM.py:484-500): minimum-distance QPSK decisions of n symbols.
 *   bits uint8 [n, 2] = (b0, b1) per symbol; hard complex64 [n] (the chosen constellation points)
 *   or NULL.  Exact ties resolve as the reference's argmin over its constellation order does.   */
GF3_API int gf3_demap(const float* symbols, int64_t n, uint8_t* bits, float* hard, void* stream);
/* transmitter.send_to_stream (OFDM.py:244-259) on symbols already in the time domain:
 *   data_time float32 [n_packets, L*(N+cp)] (CP included, no gain), sync float32 [sync_len]
 *   out float32 [n_packets*(sync_len + (2P+L)(N+cp)) + sync_len]:
 *   per packet [sync | g*(P x known) | g*data | g*(P x known)], then one trailing sync          */
GF3_API int gf3_tx_frame(const gf3_plan* plan, const float* data_time, int64_t n_packets, const float* sync,
                 int32_t sync_len, const float* known, float* out, void* stream);

/* ---- channel simulator + BER counters ---------------------- */
/* y = lfilter(taps, 1, x) + sigma * N(0,1) (Philox-4x32-10 counter RNG, Box-Muller).
 *   taps [n_streams, n_taps] float32 (n_taps <= 64), sigma [n_streams] float32            */
GF3_API int gf3_channel_sim(const float* x, int64_t x_stride, int64_t n_streams, int64_t T,
                    const float* taps, int32_t n_taps, const float* sigma, uint64_t seed,
                    float* y, int64_t y_stride, void* stream);
/* The same with an explicit Philox stream id per row (stream_ids [n_streams] int64): the noise of a row
 * depends on (seed, id, sample index) only, not on the row's position in the batch -- what the
 * stream-sharded sweep needs to be invariant under sharding. */
GF3_API int gf3_channel_sim_ids(const float* x, int64_t x_stride, int64_t n_streams, int64_t T,
                        const float* taps, int32_t n_taps, const float* sigma, const int64_t* stream_ids,
                        uint64_t seed, float* y, int64_t y_stride, void* stream);
/* Uniform random bytes for the synthetic workloads (payload bits, filler): row r is drawn from the Philox
 * counter stream of row_ids[r] (NULL: r), so it does not depend on batching or sharding either. */
GF3_API int gf3_random_bytes(uint8_t* out, int64_t out_stride, int64_t n_rows, int64_t row_bytes,
                     const int64_t* row_ids, uint64_t seed, void* stream);
/* PCM ingest (Final System Test.ipynb:85-86 does `r = r/1.0` on the wav's native samples): plain
 * value conversion to float32 on the device, no DC removal (the reference keeps the uint8 offset of
 * 128), so narrow samples cross PCIe instead of floats.  format: 0 = uint8, 1 = int16.           */
GF3_API int gf3_pcm_to_f32(const void* pcm, int32_t format, int64_t n, float* out, void* stream);
/* counter[0] += popcount(a ^ b) over nbits (MSB-first packed), counter[1] += nbits. */
GF3_API int gf3_ber_count(const uint8_t* a, const uint8_t* b, int64_t nbits, uint64_t* counter,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GF3_B200_H */
