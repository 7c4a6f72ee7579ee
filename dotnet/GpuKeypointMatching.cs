// GpuKeypointMatching.cs -- the P/Invoke engine behind dotnet/KeypointMatching.cs (the file that REPLACES
// dotnet_src/ImageProcessing/KeypointMatching.cs:8-69).  It marshals two List<Keypoint> into packed little-endian
// descriptor rows, calls pgm_match_hamming_greedy in libpgmatch.so and rebuilds the List<KeypointPair> from the returned
// indices.  It is NOT itself a drop-in: the reference's KeypointMatching is a concrete, non-virtual class with no interface
// (SURVEY D2), TestService takes a `KeypointMatching` parameter (TestService.cs:34), so a differently named or derived class
// cannot be resolved in its place -- hence the same-named replacement class next to this file.
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no .NET SDK (see DESIGN.md section 6).  The same C ABI is
// exercised from Python (ctypes) by tests/test_gpu_parity.py, which mirrors this marshalling step for step.
//
// Integration (INTEGRATION.md):
//   1. copy this file and dotnet/KeypointMatching.cs into dotnet_src/ImageProcessing/, the latter OVER the reference's
//      KeypointMatching.cs (add <AllowUnsafeBlocks>true</AllowUnsafeBlocks> to ImageProcessing.csproj);
//   2. nothing else changes: Program.cs:55, TestService.cs:34,43,96 compile and run as they are;
//   3. ship libpgmatch.so next to the executable (or on LD_LIBRARY_PATH).
using System.Numerics;
using System.Runtime.InteropServices;
using ImageProcessing.Abstractions;

namespace ImageProcessing;

public sealed class GpuKeypointMatching : IDisposable
{
    private const string Lib = "pgmatch";          // libpgmatch.so
    private const uint PGM_FLAG_REFERENCE_COMPAT_TAIL = 0x1;
    private const int PGM_OK = 0, PGM_E_EMPTY_TRAIN = -5, PGM_E_NO_DEVICE = -7;

    [DllImport(Lib)] private static extern int pgm_create(int deviceOrdinal, out IntPtr handle);
    [DllImport(Lib)] private static extern int pgm_destroy(IntPtr handle);
    [DllImport(Lib)] private static extern IntPtr pgm_last_error(IntPtr handle);
    [DllImport(Lib)] private static extern IntPtr pgm_status_string(int status);
    [DllImport(Lib)] private static extern unsafe int pgm_match_hamming_greedy(
        IntPtr handle, byte* q, int n1, byte* t, int n2, int descBits, int strideBytes,
        int* outQi, int* outTj, int* outDist, int capacity, out int outCount, uint flags);

    private readonly IntPtr _handle;
    private readonly int _descBits;

    /// <param name="descBits">KeypointDetectionOptions.NumGaussianPairs (appsettings.json:23), 256 by default.</param>
    public GpuKeypointMatching(int deviceOrdinal = 0, int descBits = 256)
    {
        var rc = pgm_create(deviceOrdinal, out _handle);
        if (rc != PGM_OK)   // there is no CPU fallback by design
            throw new InvalidOperationException(
                $"libpgmatch: {Marshal.PtrToStringAnsi(pgm_status_string(rc))}");
        _descBits = descBits;
    }

    /// Same contract as KeypointMatching.MatchKeypoints (KeypointMatching.cs:14): returns
    /// keypoints1.Count pairs in the reference's order, Keypoint1/Keypoint2 reference-equal to the
    /// inputs, including the (keypoints1[0], keypoints2[0], int.MaxValue) tail when
    /// keypoints1.Count > keypoints2.Count, and ArgumentOutOfRangeException when keypoints2 is empty
    /// and keypoints1 is not (KeypointMatching.cs:61).
    public unsafe List<KeypointPair> MatchKeypoints(List<Keypoint> keypoints1, List<Keypoint> keypoints2)
    {
        int n1 = keypoints1.Count, n2 = keypoints2.Count;
        int bits = _descBits;
        foreach (var k in keypoints1) bits = Math.Max(bits, (int)k.BriefDescriptor.GetBitLength());
        foreach (var k in keypoints2) bits = Math.Max(bits, (int)k.BriefDescriptor.GetBitLength());
        int stride = (bits + 127) / 128 * 16;
        var q = Pack(keypoints1, stride);
        var t = Pack(keypoints2, stride);
        var qi = new int[Math.Max(n1, 1)];
        var tj = new int[Math.Max(n1, 1)];
        var dd = new int[Math.Max(n1, 1)];
        int rc, count;
        fixed (byte* pq = q, pt = t)
        fixed (int* pqi = qi, ptj = tj, pdd = dd)
            rc = pgm_match_hamming_greedy(_handle, pq, n1, pt, n2, bits, stride, pqi, ptj, pdd, n1, out count,
                                          PGM_FLAG_REFERENCE_COMPAT_TAIL);
        if (rc == PGM_E_EMPTY_TRAIN)
            throw new ArgumentOutOfRangeException("index");      // what keypoints2[0] throws upstream
        if (rc != PGM_OK)
            throw new InvalidOperationException($"libpgmatch: {Marshal.PtrToStringAnsi(pgm_last_error(_handle))}");

        var pairs = new List<KeypointPair>(count);
        for (int k = 0; k < count; k++)
            pairs.Add(new KeypointPair
            {
                Distance = dd[k],
                Keypoint1 = keypoints1[qi[k]],     // KeypointMatching.cs:57-62: references into the inputs
                Keypoint2 = keypoints2[tj[k]],
            });
        return pairs;
    }

    // BigInteger.ToByteArray(isUnsigned: true) drops leading zero bytes, so every descriptor is
    // copied into its own zero-filled stride-byte slot (little-endian).
    private static byte[] Pack(List<Keypoint> keypoints, int stride)
    {
        var buf = new byte[Math.Max(keypoints.Count, 1) * stride];
        for (int k = 0; k < keypoints.Count; k++)
        {
            BigInteger d = keypoints[k].BriefDescriptor;
            if (d.Sign < 0) throw new ArgumentException("BriefDescriptor must be non-negative");
            d.TryWriteBytes(buf.AsSpan(k * stride, stride), out _, isUnsigned: true, isBigEndian: false);
        }
        return buf;
    }

    public void Dispose() => pgm_destroy(_handle);
}
