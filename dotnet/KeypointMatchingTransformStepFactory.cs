// KeypointMatchingTransformStepFactory.cs -- the matcher as a TPL-Dataflow step, in the style of
// dotnet_src/ImageProcessing/PipelinesV3/Factories/RedundantKeypointEliminatorTransformStepFactory.cs:29-41
// (interface: ImageProcessing.Abstractions/PipelinesV3/ITransformStepFactory.cs:5-8).
//
// The reference calls the matcher imperatively after its pipeline has drained (TestService.cs:89-96).  This step does the
// same work inside the pipeline for a frame sequence (BASELINE configs[2]): every record that arrives is matched against
// the record that arrived before it (consecutive frames), using the de-noised keypoints the preceding
// RedundantKeypointEliminator step stored.  MetadataStore has no slot for match lists (MetadataVariant.cs:3-11), so the
// step keeps them itself, keyed by the two record guids.  A TransformBlock runs one item at a time by default
// (MaxDegreeOfParallelism = 1, as everywhere in TestService.BuildKeypointDetectorPipeline), which this step relies on.
//
// Place in dotnet_src/ImageProcessing/PipelinesV3/Factories/ and link after the eliminator block:
//     var matchStep = new KeypointMatchingTransformStepFactory(_metadataStore, _keypointMatching).GetAndInitTransformBlock();
//     redundantKeypointEliminatorBlock.LinkTo(matchStep, linkOptions);
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no .NET SDK (DESIGN.md section 6).
using System.Collections.Concurrent;
using System.Threading.Tasks.Dataflow;
using ImageProcessing.Abstractions;
using ImageProcessing.Abstractions.PipelinesV3;
using ImageProcessing.PipelinesV3.DTOs;
using Microsoft.Extensions.Logging;
using PhotogrammetryStore;

namespace ImageProcessing.PipelinesV3.Factories;

public class KeypointMatchingTransformStepFactory : ITransformStepFactory<MetadataStoreRecord, MetadataStoreRecord>
{
    private readonly MetadataStore _metadataStore;
    private readonly KeypointMatching _keypointMatching;
    private readonly ILogger<KeypointMatchingTransformStepFactory>? _logger;
    private MetadataStoreRecord? _previous;

    /// (previous record, this record) -> the reference-ordered match list of the two frames.
    public ConcurrentDictionary<(Guid Previous, Guid Current), List<KeypointPair>> Matches { get; } = new();

    public KeypointMatchingTransformStepFactory(MetadataStore metadataStore, KeypointMatching keypointMatching,
        ILogger<KeypointMatchingTransformStepFactory>? logger = null)
    {
        _metadataStore = metadataStore;
        _keypointMatching = keypointMatching;
        _logger = logger;
    }

    public void Initialize()
    {
        _previous = null;
    }

    public TransformBlock<MetadataStoreRecord, MetadataStoreRecord> GetTransformBlock()
    {
        return new TransformBlock<MetadataStoreRecord, MetadataStoreRecord>(record =>
        {
            if (_previous is { } previous)
            {
                _logger?.LogInformation("Matching Keypoints");
                var kp1 = _metadataStore.FetchDenoisedKeypoints(previous.RecordGuid);
                var kp2 = _metadataStore.FetchDenoisedKeypoints(record.RecordGuid);
                var keypointPairs = _keypointMatching.MatchKeypoints(kp1, kp2);
                Matches[(previous.RecordGuid, record.RecordGuid)] = keypointPairs;
                _logger?.LogInformation("Matched keypoints. Found {Ct} pairs", keypointPairs.Count);
            }

            _previous = record;
            return record;
        });
    }

    public TransformBlock<MetadataStoreRecord, MetadataStoreRecord> GetAndInitTransformBlock()
    {
        Initialize();
        return GetTransformBlock();
    }
}
