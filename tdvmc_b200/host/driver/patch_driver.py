#!/usr/bin/env python3
"""Writes the GPU-bound copy of the reference driver:

    patch_driver.py <reference>/src/TDVMC.cpp <out>/TDVMC_gpu.cpp

The reference's src/TDVMC.cpp is read where it lies and every edit below is applied to an in-memory copy; the result
goes to the build directory (git-ignored - no reference source enters this repository).  The edits are exactly the
binding INTEGRATION.md section 2 describes: two #include lines for TDVMC_gpu_hooks.h (our code) and one-line calls at
the driver's per-rank sampling entry points.  Every anchor must match exactly once, otherwise the script fails - a
changed upstream driver is an error, never a silent no-op.

Reference lines (mathiasgartner/TDVMC, src/TDVMC.cpp) each edit sits at are given with the edit.
"""
import re
import sys

INCLUDE_DECL = '#define TDVMC_GPU_HOOKS_DECL\n#include "TDVMC_gpu_hooks.h"\n'
INCLUDE_IMPL = '#define TDVMC_GPU_HOOKS_IMPL\n#include "TDVMC_gpu_hooks.h"\n'

# (description, anchor regex, replacement template using \g<0> for the matched anchor)
EDITS = [
    (":180 after the driver's globals: declarations of the GPU hooks and the new config globals",
     r"vector<SimulationStepData> previousStepData;\n",
     r"\g<0>" + INCLUDE_DECL),
    (":348 RegisterAllConfigItems: GPU_WALKERS, GPU_SEED, GPU_DEVICE_SOLVE registered like the reference's own items",
     r'\tconfigItems\.push_back\(ConfigItem\("PARAM_PHII", &PARAM_PHII, ConfigItemType::DOUBLE\)\);\n',
     r"\g<0>\tGpuRegisterConfigItems();\n"),
    (":975 UpdateSamplesConsecutive: the stored samples live on the device",
     r"void UpdateSamplesConsecutive\(int nrOfSamplesToUpdate, [^)]*\)\n\{\n",
     r"\g<0>\tif (GpuUpdateSamplesConsecutive(nrOfSamplesToUpdate, uR, uI, phiR, phiI))\n\t{\n\t\treturn;\n\t}\n"),
    (":1152 ParallelUpdateExpectationValues: estimator pass + the seven ReduceToAverage calls",
     r"void ParallelUpdateExpectationValues\(vector<vector<double> >& R, [^)]*\)\n\{\n",
     r"\g<0>\tif (GpuParallelUpdateExpectationValues(uR, uI, phiR, phiI, intermediateStep))\n\t{\n\t\treturn;\n\t}\n"),
    (":1305 ParallelUpdateExpectationValuesForGivenSamples",
     r"void ParallelUpdateExpectationValuesForGivenSamples\(vector<ICorrelatedSamplingData\*>& samples, [^)]*\)\n\{\n",
     r"\g<0>\tif (GpuParallelUpdateExpectationValuesForGivenSamples(uR, uI, phiR, phiI))\n\t{\n\t\treturn;\n\t}\n"),
    (":1438 ParallelCalculateAdditionalSystemProperties: g(r), S(k)",
     r"void ParallelCalculateAdditionalSystemProperties\(vector<vector<double> >& R, [^)]*\)\n\{\n",
     r"\g<0>\tif (GpuParallelCalculateAdditionalSystemProperties(uR, uI, phiR, phiI))\n\t{\n\t\treturn;\n\t}\n"),
    (":1834 CalculateNextParametersEuler: optional device solve (GPU_DEVICE_SOLVE = 1, Cholesky branch)",
     r"void CalculateNextParametersEuler\(double dt, [^)]*\)\n\{\n",
     r"\g<0>\tif (GpuCalculateNextParametersEuler(dt, uR, uI, phiR, phiI))\n\t{\n\t\treturn;\n\t}\n"),
    (":2569 AlignCoordinates: the device walkers are wrapped into the first cell with the driver's R",
     r"void AlignCoordinates\(vector<vector<double> >& R\)\n\{\n",
     r"\g<0>\tGpuAlignCoordinates();\n"),
    (":3004 before mainMPI: definitions of the hooks (they use the driver's own helpers)",
     r"int mainMPI\(int argc, char\*\* argv\)\n\{\n",
     INCLUDE_IMPL + r"\g<0>"),
    (":3132-3133 after sys->InitSystem(); PostSystemInit(); in mainMPI: create the device ensemble from the system's tables",
     r"(?<=\n)\tsys->InitSystem\(\);\n\tPostSystemInit\(\);\n(?=\t//Write grid files for observables)",
     r"\g<0>\tGpuInit(R);\n"),
    (":3413-3416 MC_VERY_FIRST_NINITIALIZATIONSTEPS of the time-evolution driver run on the device walkers",
     r"(\tsys->CalculateWavefunction\(R, uR, uI, phiR, phiI\);\n)(\tfor \(int i = 0; i < )(MC_VERY_FIRST_NINITIALIZATIONSTEPS)(; i\+\+\)\n\t\{\n\t\tDoMetropolisStep\(R, uR, uI, phiR, phiI\);\n\t\}\n\tsys->CalculateWavefunction\(R, uR, uI, phiR, phiI\);\n\tfor \(currentTime = 0;)",
     r"\1\tGpuVeryFirstInitialization(uR, uI, phiR, phiI);\n\2(gpu ? 0 : \3)\4"),
    (":3433-3434 start of a time step: acceptance counters restart on the device as on the host (:3428-3429)",
     r"(?<=\n)\t\tsys->SetTime\(currentTime\);\n\t\tsys->SetStep\(step\);\n\t\ttimes\.push_back\(currentTime\);\n",
     r"\g<0>\t\tGpuBeginTimeStep();\n"),
    (":3763 NormalizeWavefunction takes the exponent of the device's last sample",
     r"\t\t\t\t\tNormalizeWavefunction\(sys->GetExponent\(\), &phiR\);\n",
     "\t\t\t\t\tNormalizeWavefunction(GpuExponentOr(sys->GetExponent()), &phiR);\n"),
    (":4011 end of run: the coordinates written for the next run are the first local walker's",
     r"\t// Write config file for successive simulations\n",
     r"\g<0>\tGpuCopyWalkerToDriver(R);\n"),
    (":4050 end of run: release the device",
     r"\tadditionalObservablesMean\.Destroy\(\);\n\n\t//Log\(\"finalize \.\.\.\"\);\n",
     r"\tGpuShutdown();\n\g<0>"),
]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    text = open(src).read()
    for what, anchor, repl in EDITS:
        n = len(re.findall(anchor, text))
        if n != 1:
            sys.exit(f"patch_driver: anchor for edit '{what}' matched {n} times in {src} (expected exactly 1)")
        text = re.sub(anchor, repl, text, count=1)
    header = ("// GENERATED by tdvmc_b200/host/driver/patch_driver.py from the reference's src/TDVMC.cpp - do not edit, do not commit.\n"
              f"// {len(EDITS)} edits: see patch_driver.py for each one and the reference line it sits at.\n")
    open(dst, "w").write(header + text)
    print(f"patch_driver: {len(EDITS)} edits applied -> {dst}")


if __name__ == "__main__":
    main()
