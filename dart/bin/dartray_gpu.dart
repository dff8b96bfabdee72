// @dart=2.9
// The CLI of bin/dartray.dart with its render call restored (the reference's is commented out,
// bin/dartray.dart:43-51, and lib/dartray_io/render_manager.dart has no render()): parse the .pbrt scene with the
// unchanged Dart front end, render through GpuOrDartRenderer (see dart/README.md for the two-line patch of
// DartRay._makeRenderer), write the PNG.  REVIEWED, NOT RUN: no Dart SDK in the build image.
import 'dart:io';
import 'package:args/args.dart';
import 'package:dartray/dartray_io.dart';
import 'package:image/image.dart';

void main(List<String> argv) {
  var parser = new ArgParser();
  parser.addOption('output', abbr: 'o', defaultsTo: 'output.png');
  var args = parser.parse(argv);
  if (args.rest.isEmpty) {
    print('Usage: dartray_gpu [options] <scene.pbrt>');
    print(parser.usage);
    return;
  }
  String out = args['output'];
  String scene = args.rest[0];
  Stopwatch timer = new Stopwatch()..start();
  // what lib/dartray_web/render_manager.dart:93 does in the browser: DartRay.renderScene on the manager's loader
  var manager = new RenderManager(new File(scene).parent.path);
  new DartRay(manager).renderScene(scene).then((OutputImage output) {
    timer.stop();
    LogInfo('RENDER FINISHED: ${timer.elapsed}');
    if (output != null) {
      Image image = output.toImage();
      new File(out).writeAsBytesSync(encodeNamedImage(image, out));
    }
  });
}
