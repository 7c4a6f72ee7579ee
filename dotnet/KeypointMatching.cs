// KeypointMatching.cs -- REPLACES dotnet_src/ImageProcessing/KeypointMatching.cs.
//
// Same namespace, same class name, same public parameterless constructor (KeypointMatching.cs:10-12), same
//     List<KeypointPair> MatchKeypoints(List<Keypoint> keypoints1, List<Keypoint> keypoints2)      (KeypointMatching.cs:14)
// so every caller compiles untouched: the DI registration `services.AddSingleton<KeypointMatching>()`
// (Photogrammetry/Program.cs:55), the constructor injection into TestService (TestService.cs:34,43) and the call at
// TestService.cs:96.  The reference class is concrete, non-virtual and has no interface (SURVEY D2), so replacing the file
// is the only drop-in that needs no change elsewhere; the work itself is done by GpuKeypointMatching (P/Invoke into
// libpgmatch.so, next file), created on first use and shared by the process.
//
// Behaviour kept (tests/test_gpu_parity.py checks each point through the same C ABI from Python):
//   * keypoints1.Count pairs, in the order the reference's repeated global-argmin scan emits them (:38-66);
//   * Keypoint1 / Keypoint2 are the caller's own objects (:57-62);
//   * the (keypoints1[0], keypoints2[0], int.MaxValue) tail when keypoints1.Count > keypoints2.Count (:38-42);
//   * ArgumentOutOfRangeException when keypoints2 is empty and keypoints1 is not (:61).
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no .NET SDK (DESIGN.md section 6).
using ImageProcessing.Abstractions;

namespace ImageProcessing;

public class KeypointMatching
{
    // one handle (= one GPU stream + scratch) per process; calls on it are serialised inside the library
    private static readonly Lazy<GpuKeypointMatching> Gpu =
        new(() => new GpuKeypointMatching(DeviceOrdinalFromEnvironment()), LazyThreadSafetyMode.ExecutionAndPublication);

    public KeypointMatching()
    {
    }

    public List<KeypointPair> MatchKeypoints(List<Keypoint> keypoints1, List<Keypoint> keypoints2)
    {
        return Gpu.Value.MatchKeypoints(keypoints1, keypoints2);
    }

    // PGMATCH_DEVICE selects the GPU (default 0); there is deliberately no CPU fallback: without a B200 the first call throws.
    private static int DeviceOrdinalFromEnvironment()
    {
        return int.TryParse(Environment.GetEnvironmentVariable("PGMATCH_DEVICE"), out var d) ? d : 0;
    }
}
